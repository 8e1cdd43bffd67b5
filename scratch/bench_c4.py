"""BASELINE config 4, one GPU's share: G galaxies of 1e3..1e5 particles, SEDs + images in batched calls."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from synthobs_b200 import synthetic
from synthobs_b200.sed import models
from synthobs_b200.morph import images
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
model, F, cosmo = bench.setup_model()
b = synthetic.galaxy_batch(G, n_min=1000, n_max=100000, seed=5)
n = b['Masses'].size
print('G=%d galaxies, %d particles (%.1f MB/array)' % (G, n, n * 8 / 1e6))
dev = torch.device('cuda')
d = {k: torch.from_numpy(b[k]).to(dev) for k in ('Masses', 'Ages', 'Metallicities', 'tauVs_ISM', 'tauVs_BC', 'X', 'Y')}
off = torch.from_numpy(b['offsets']).to(dev)
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r
ms, r = timeit(lambda: models.generate_Fnu_batch(model, F, d['Masses'], d['Ages'], d['Metallicities'], off, tauVs_ISM=d['tauVs_ISM'], tauVs_BC=d['tauVs_BC']))
print('generate_Fnu_batch (10 filters): %.3f ms -> %.3e galaxy SEDs x filters/s, %.3e particles/s' % (ms, G * 10 / ms * 1e3, n / ms * 1e3))
ms, r = timeit(lambda: models.generate_SED_batch(model, d['Masses'], d['Ages'], d['Metallicities'], off, tauVs_ISM=d['tauVs_ISM'], tauVs_BC=d['tauVs_BC']), reps=2)
print('generate_SED_batch (5 spectra x 9000): %.2f ms -> %.1f galaxy SEDs/s' % (ms, G / ms * 1e3))
Lnu = models.generate_Fnu_arrays(model, F, d['Masses'], d['Ages'], d['Metallicities'], tauVs_ISM=d['tauVs_ISM'], tauVs_BC=d['tauVs_BC'])
L8 = torch.stack([Lnu[f] for f in F['filters'][:8]]).to(torch.float64)
for C, L in ((1, L8[0]), (8, L8)):
    ms, r = timeit(lambda: images.core_batch(d['X'], d['Y'], L, off, resolution=0.1, ndim=100, smoothing=('adaptive', 16)))
    print('core_batch %d channel(s), 100x100, adaptive 16: %.2f ms -> %.3e particles/s, %.1f galaxy images/s ; deposits %.3e -> %.3e/s'
          % (C, ms, n / ms * 1e3, G * C / ms * 1e3, r.info['deposits'] * C, r.info['deposits'] * C / ms * 1e3))
# the per-object loop the reference API implies, on a sample
t0 = time.time(); m = 40
for g in range(m):
    s, e = int(b['offsets'][g]), int(b['offsets'][g + 1])
    images.core(d['X'][s:e], d['Y'][s:e], L8[0][s:e], resolution=0.1, ndim=100, smoothing=('adaptive', 16))
torch.cuda.synchronize()
print('loop of core() over objects: %.2f ms per object (sample of %d)' % ((time.time() - t0) / m * 1e3, m))
