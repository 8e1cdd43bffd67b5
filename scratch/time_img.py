import sys; sys.path.insert(0, '.')
import numpy as np, torch
import bench
from synthobs_b200 import synthetic, _cabi
from synthobs_b200.morph import images
from ctypes import c_double, c_int64
dev = torch.device('cuda')
p = synthetic.particles(bench.IMG_PARTICLES, seed=200, scale_kpc=3.0)
filters = ['JWST.NIRCAM.B%d' % i for i in range(8)]
for f in filters: images.filter_info[f] = {'pixel_scale': 0.031}
cosmo = synthetic.FixedCosmology()
obs = [images.observed(f, cosmo, 8.0, 512 * 0.031 + 1e-9, smoothing=('adaptive', 8.), PSF=synthetic.GaussianPSF(1.5 + 0.2 * i), super_sampling=2) for i, f in enumerate(filters)]
Xd = torch.from_numpy(p['X']).to(dev); Yd = torch.from_numpy(p['Y']).to(dev)
L = torch.from_numpy(np.stack([p['Masses']] * 8)).to(dev)
o0 = obs[0]
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(4):
    t = [ev()]
    X = Xd - images.median_device(Xd); Y = Yd - images.median_device(Yd); t.append(ev())
    width, G = images.pixel_grid(o0.resolution, o0.ndim_super); Gd = images.dv.to_dev(G)
    ix, iy = images.ngp_index_device(X, Y, Gd, o0.ndim_super, width / 2.); t.append(ev())
    dk = images.knn_device(X, Y, ix, 8, width / 2.); t.append(ev())
    r = images.deposit_device(X, Y, L, o0.resolution, o0.ndim_super, adaptive_k=8, want_ngp=False, dk=dk); t.append(ev())
    if it == 0:
        spec = images.KernelSpectrum(np.stack([o.psf_kernel() for o in obs]), o0.ndim_super, o0.ndim_super)
        t[-1] = ev()
    conv = images.convolve_fft_device(r['data'], spec); t.append(ev())
    a = images.rebin_device(r['data'], o0.ndim); b = images.rebin_device(conv, o0.ndim); t.append(ev())
    torch.cuda.synchronize()
names = ['median+recentre', 'grid+ngp index', 'knn', 'plan+deposit', 'psf fft', 'rebin']
print(' | '.join('%s %.3f' % (n, t[i].elapsed_time(t[i + 1])) for i, n in enumerate(names)), '| total %.3f ms' % t[0].elapsed_time(t[-1]))
