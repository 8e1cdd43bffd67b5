"""Short driver for ncu: one or two launches of every hot kernel at bench sizes."""
import sys; sys.path.insert(0, '.')
import numpy as np, torch
import bench
from synthobs_b200 import _cabi, synthetic
from synthobs_b200.sed import models
from synthobs_b200.morph import images
dev = torch.device('cuda', 0)
model, F, cosmo = bench.setup_model()
b = synthetic.galaxy_batch(256, fixed_n=100000, seed=100)
keys = ('Masses', 'Ages', 'Metallicities', 'tauVs_ISM', 'tauVs_BC')
d = {k: torch.from_numpy(b[k]).to(dev) for k in keys}; off = torch.from_numpy(b['offsets']).to(dev)
for _ in range(1):
    models.generate_Fnu_batch(model, F, d['Masses'], d['Ages'], d['Metallicities'], off, tauVs_ISM=d['tauVs_ISM'], tauVs_BC=d['tauVs_BC'])
torch.cuda.synchronize()
n32 = 32 * 100000
off32 = off[:33].clone()
for _ in range(1):
    models.sed_spectra_device(model, d['Masses'][:n32], d['Ages'][:n32], d['Metallicities'][:n32], d['tauVs_ISM'][:n32], d['tauVs_BC'][:n32], 0.0, 7.0, offsets=off32)
torch.cuda.synchronize()
ops = models._spectral_operands(model)
A = torch.rand((20000, ops['ld']), device=dev); Ah, Al = models.split_tf32(A); Bh, Bl = ops['BT_split']
C = torch.empty((20000, model.lam.size), device=dev)
for _ in range(1):
    _cabi.check(_cabi.lib.so_gemm_3xtf32(_cabi.ptr(Ah), _cabi.ptr(Al), A.stride(0), _cabi.ptr(Bh), _cabi.ptr(Bl), Bh.stride(0), _cabi.ptr(C), C.stride(0), 20000, model.lam.size, 2 * model.na * model.nz, _cabi.stream()), 'gemm')
torch.cuda.synchronize()
p = synthetic.particles(bench.IMG_PARTICLES, seed=200, scale_kpc=3.0)
filters = ['JWST.NIRCAM.B%d' % i for i in range(8)]
for f in filters: images.filter_info[f] = {'pixel_scale': 0.031}
obs = [images.observed(f, cosmo, 8.0, 512 * 0.031 + 1e-9, smoothing=('adaptive', 8.), PSF=synthetic.GaussianPSF(1.5 + 0.2 * i), super_sampling=2) for i, f in enumerate(filters)]
X = torch.from_numpy(p['X'] - np.median(p['X'])).to(dev); Y = torch.from_numpy(p['Y'] - np.median(p['Y'])).to(dev)
L = torch.from_numpy(np.stack([p['Masses']] * 8)).to(dev)
Xr, Yr = torch.from_numpy(p['X']).to(dev), torch.from_numpy(p['Y']).to(dev)
for _ in range(1):
    images.median_device_pair(Xr, Yr)
    images.particle_device_multi(obs, X, Y, L)
torch.cuda.synchronize()
# the reference-named call (per-filter recentring chain + the same pipeline)
PSFs = {f: synthetic.GaussianPSF(1.5 + 0.2 * i) for i, f in enumerate(filters)}
images.particle(Xr.clone(), Yr.clone(), {f: L[i] for i, f in enumerate(filters)}, filters, cosmo, 8.0, 512 * 0.031 + 1e-9,
                smoothing=('adaptive', 8.), PSFs=PSFs, super_sampling=2)
torch.cuda.synchronize()
print('done')
