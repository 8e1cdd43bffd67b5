import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from synthobs_b200 import synthetic
from synthobs_b200.morph import images
N = 1000000
p = synthetic.particles(N, seed=200, scale_kpc=3.0)
X = torch.from_numpy(p['X'] - np.median(p['X'])).cuda(); Y = torch.from_numpy(p['Y'] - np.median(p['Y'])).cuda()
rng = np.random.default_rng(0)
ndim = 1024
res = 0.031 / 0.2033 / 2
for C in (1, 8):
    L = torch.from_numpy(np.stack([p['Masses'] * (0.5 + rng.random(N)) for _ in range(C)])).cuda()
    for k in (8, 32, 128, 512):
        dk = images.knn_device(X, Y, images.ngp_index_device(X, Y, torch.from_numpy(images.pixel_grid(res, ndim)[1]).cuda(), ndim, ndim * res / 2)[0], k, ndim * res / 2)
        def run():
            return images.deposit_device(X, Y, L, res, ndim, adaptive_k=k, want_ngp=False, dk=dk)
        r = run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): r = run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        dep = r['stats']['deposits'] * C
        print('C=%d k=%4d: plan+deposit %.3f ms  pairs %.3g  deposits %.3g  -> %.3g deposits/s (%.1f%% of FP32 FMA peak 3.6e13)' % (C, k, ms, r['stats']['pairs'], dep, dep / ms * 1e3, dep / ms * 1e3 / 3.6e13 * 100))
