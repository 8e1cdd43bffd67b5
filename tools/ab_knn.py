"""A/B timing of the k-NN search structures (SO_KNN_ALGO is read once per process: run one process per setting).

    SO_KNN_ALGO=quadtree python tools/ab_knn.py ; python tools/ab_knn.py

Prints, per case, the whole so_knn_kth_distance_ex call by CUDA events (median of 10) and the per-kernel times of the
library's own profiler, plus a checksum of the distances (must agree between the settings bit for bit).
"""
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch
from synthobs_b200 import _cabi, synthetic
from synthobs_b200.morph import images

dev = torch.device('cuda', 0)


def case(name, n, k, width, ndim, floor, scale=3.0, batch=0):
    p = synthetic.particles(n, seed=200, scale_kpc=scale)
    X = torch.from_numpy(p['X'] - np.median(p['X'])).to(dev)
    Y = torch.from_numpy(p['Y'] - np.median(p['Y'])).to(dev)
    width, _, G = images.pixel_grid_device(width / ndim, ndim)
    ix, iy = images.ngp_index_device(X, Y, G, ndim, width / 2.)
    img, n_img = None, 1
    if batch:
        img = (torch.arange(n, device=dev) * batch // n).to(torch.int32)
        n_img = batch
    for _ in range(3):
        dk = images.knn_device(X, Y, ix, k, width / 2., img=img, n_img=n_img, floor=floor)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dk = images.knn_device(X, Y, ix, k, width / 2., img=img, n_img=n_img, floor=floor)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    _cabi.prof_enable(True)
    for _ in range(3):
        images.knn_device(X, Y, ix, k, width / 2., img=img, n_img=n_img, floor=floor)
    torch.cuda.synchronize()
    rep = _cabi.prof_report(); _cabi.prof_enable(False)
    h = dk.cpu().numpy()
    fin = np.isfinite(h)
    print('%-34s n=%-9d k=%-3d call %.3f ms   sum %.17g  max %.17g  ninf %d' % (
        name, n, k, float(np.median(ts)), h[fin].sum(), h[fin].max(), int((~fin).sum())), flush=True)
    print('      ' + '  '.join('%s %.3f' % (kk.replace('_kernel', ''), ms / c) for kk, (c, ms) in rep.items()), flush=True)


print('SO_KNN_ALGO =', os.environ.get('SO_KNN_ALGO', '(default)'))
W = 512 * 0.031 + 1e-9
if '--sizes' in sys.argv:
    for n in (10000, 30000, 100000, 300000, 1000000):
        case('clustered k8', n, 8, W, 1024, images.FWHM_FLOOR)
        case('clustered k16', n, 16, W, 1024, images.FWHM_FLOOR)
    sys.exit(0)
case('config2 exact', 1000000, 8, W, 1024, 0.0)
case('config2 floor', 1000000, 8, W, 1024, images.FWHM_FLOOR)
case('config2 k16 floor', 1000000, 16, W, 1024, images.FWHM_FLOOR)
case('config0 1e4 k16', 10000, 16, 10.0, 100, images.FWHM_FLOOR, scale=1.5)
case('batch 1000 x 2e4 k16', 20000000, 16, 10.0, 100, images.FWHM_FLOOR, scale=1.5, batch=1000)
if '--big' in sys.argv:
    case('config5 1e8 floor', 100000000, 8, 8192 * 0.004, 8192, images.FWHM_FLOOR, scale=4.0)
