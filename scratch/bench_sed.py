import sys, time, io, contextlib
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from synthobs_b200 import synthetic
from synthobs_b200.sed import models
model, F, cosmo = bench.setup_model(n_lam=600)
G = 256
b = synthetic.galaxy_batch(G, fixed_n=100000, seed=100)
d = {k: torch.from_numpy(b[k]).cuda() for k in ('Masses', 'Ages', 'Metallicities', 'tauVs_ISM', 'tauVs_BC')}
off = torch.from_numpy(b['offsets']).cuda()
def step(**kw):
    return models.generate_Fnu_batch(model, F, d['Masses'], d['Ages'], d['Metallicities'], off, tauVs_ISM=d['tauVs_ISM'], tauVs_BC=d['tauVs_BC'], **kw)
for name, kw in (('fesc0', {}), ('fesc0.3', {'fesc': 0.3})):
    for _ in range(3): step(**kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): step(**kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(name, 'ms/step %.4f  GB/s %.0f  frac %.3f' % (ms, 40 * d['Masses'].numel() / ms / 1e6, 40 * d['Masses'].numel() / ms / 1e6 / 6551.7))
# single galaxy (no segments) variant and per-particle outputs
for _ in range(3): models.generate_Fnu_arrays(model, F, d['Masses'], d['Ages'], d['Metallicities'], tauVs_ISM=d['tauVs_ISM'], tauVs_BC=d['tauVs_BC'])
torch.cuda.synchronize(); t = time.time()
for _ in range(5): models.generate_Fnu_arrays(model, F, d['Masses'], d['Ages'], d['Metallicities'], tauVs_ISM=d['tauVs_ISM'], tauVs_BC=d['tauVs_BC'])
torch.cuda.synchronize(); ms = (time.time() - t) / 5 * 1e3
print('per-particle arrays: ms %.3f GB/s %.0f (80 B/particle)' % (ms, 80 * d['Masses'].numel() / ms / 1e6))
