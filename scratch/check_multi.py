"""torchrun --nproc-per-node N scratch/check_multi.py : sharded paths vs the single-rank result."""
import contextlib, io, os, sys
sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as dist
from synthobs_b200 import dist as sdist, synthetic
from synthobs_b200.sed import models
from synthobs_b200.morph import images
rank, ws = sdist.init_from_env('nccl')
grid = synthetic.ssp_grid(n_lam=200)
with contextlib.redirect_stdout(io.StringIO()):
    model = models.define_model(grid)
model.dust_ISM = model.dust_BC = ('simple', {'slope': -1.0})
F = synthetic.filter_set(grid['lam'], n_filters=10)
model.create_Lnu_grid(F)
b = synthetic.galaxy_batch(64, n_min=1000, n_max=100000, seed=3)
full = sdist.generate_Fnu_sharded(model, F, b, kind='Lnu').cpu().numpy()
ref = models.generate_Lnu_batch(model, F, b['Masses'], b['Ages'], b['Metallicities'], b['offsets'], tauVs_ISM=b['tauVs_ISM'], tauVs_BC=b['tauVs_BC'])
e1 = np.abs(full / ref - 1).max()
p = synthetic.particles(400000, seed=12)
torch.cuda.synchronize(); dist.barrier()
import time; t = time.time()
img = sdist.image_huge(p['X'], p['Y'], p['Masses'], 0.01, 2048, 8)
torch.cuda.synchronize(); dt = time.time() - t
one = images.core(p['X'], p['Y'], p['Masses'], resolution=0.01, ndim=2048, smoothing=('adaptive', 8)).data
e2 = np.abs(img.cpu().numpy() - one).max() / one.max()
print('rank %d/%d: sharded SED max rel diff %.2e ; huge image vs single-rank max diff/peak %.2e ; image_huge %.1f ms' % (rank, ws, e1, e2, dt * 1e3))
assert e1 < 1e-6 and e2 < 2e-6   # per-rank sub-batches group the FP32 partial sums differently
pp = synthetic.particles(200000, seed=13)
kw = dict(tauVs_ISM=pp['tauVs_ISM'], tauVs_BC=pp['tauVs_BC'], fesc=0.1)
sh = sdist.generate_SED_huge(model, pp['Masses'], pp['Ages'], pp['Metallicities'], **kw)
rf = models.sed_spectra_device(model, pp['Masses'], pp['Ages'], pp['Metallicities'], pp['tauVs_ISM'], pp['tauVs_BC'], 0.1, 7.0)
e3 = max(float(((sh[k] - rf[k]).abs().max() / rf[k].abs().max()).cpu()) for k in rf)
fh = sdist.generate_Fnu_huge(model, F, pp['Masses'], pp['Ages'], pp['Metallicities'], kind='Lnu', **kw)
fr = models.generate_Lnu(model, F, pp['Masses'], pp['Ages'], pp['Metallicities'], **kw)
e4 = max(abs(fh[f] / fr[f] - 1) for f in F['filters'])
print('rank %d: huge-galaxy spectra vs single rank %.2e ; fluxes %.2e' % (rank, e3, e4))
assert e3 < 2e-6 and e4 < 1e-6
b2 = synthetic.galaxy_batch(40, n_min=500, n_max=20000, seed=9)
img = sdist.core_batch_sharded(b2['X'], b2['Y'], b2['Masses'], b2['offsets'], resolution=0.1, ndim=64, smoothing=('adaptive', 8))
one = images.core_batch(b2['X'], b2['Y'], b2['Masses'], b2['offsets'], resolution=0.1, ndim=64, smoothing=('adaptive', 8)).data
e5 = np.abs(img.cpu().numpy() - one).max() / one.max()
sp = sdist.generate_SED_sharded(model, b2, fesc=0.1)
rf2 = models.generate_SED_batch(model, b2['Masses'], b2['Ages'], b2['Metallicities'], b2['offsets'], tauVs_ISM=b2['tauVs_ISM'], tauVs_BC=b2['tauVs_BC'], fesc=0.1)
e6 = max(np.abs(sp[k].cpu().numpy() - rf2[k]).max() / np.abs(rf2[k]).max() for k in rf2)
print('rank %d: sharded batch images vs one rank %.2e ; sharded spectra %.2e' % (rank, e5, e6))
assert e5 < 3e-6 and e6 < 2e-6
dist.destroy_process_group()
