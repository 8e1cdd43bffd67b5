import sys, os
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from synthobs_b200 import synthetic
from synthobs_b200.sed import models
model, F, cosmo = bench.setup_model()
G = 32
b = synthetic.galaxy_batch(G, fixed_n=100000, seed=300)
d = {k: torch.from_numpy(b[k]).cuda() for k in ('Masses', 'Ages', 'Metallicities', 'tauVs_ISM', 'tauVs_BC')}
off = torch.from_numpy(b['offsets']).cuda()
def step():
    return models.sed_spectra_device(model, d['Masses'], d['Ages'], d['Metallicities'], d['tauVs_ISM'], d['tauVs_BC'], 0.0, 7.0, offsets=off)
for _ in range(2): o = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): o = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
nyoung = 0
print('hybrid=%s spectra step ms %.3f  particle-lambda/s %.3e' % (os.environ.get('SO_DUST_HYBRID', '1'), ms, G * 1e5 * 9000 / ms * 1e3))
np.save('gpurun_out/dust_total_%s.npy' % os.environ.get('SO_DUST_HYBRID', '1'), o['total'].cpu().numpy())
