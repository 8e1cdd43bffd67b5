import sys
sys.path.insert(0, '.')
import numpy as np, torch
from synthobs_b200 import synthetic
from synthobs_b200.morph import images
N = 1000000
p = synthetic.particles(N, seed=200, scale_kpc=3.0)
X = torch.from_numpy(p['X'] - np.median(p['X'])).cuda(); Y = torch.from_numpy(p['Y'] - np.median(p['Y'])).cuda()
ndim = 1024; res = 0.031 / 0.2033 / 2
G = torch.from_numpy(images.pixel_grid(res, ndim)[1]).cuda()
ix, iy = images.ngp_index_device(X, Y, G, ndim, ndim * res / 2)
for k in (8, 16):
    for fl in (0.0, 0.01):
        for _ in range(2): dk = images.knn_device(X, Y, ix, k, ndim * res / 2, floor=fl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): dk = images.knn_device(X, Y, ix, k, ndim * res / 2, floor=fl)
        e1.record(); torch.cuda.synchronize()
        print('k=%d floor=%g knn total %.3f ms' % (k, fl, e0.elapsed_time(e1) / 5))
