"""SO_KNN_STATS=1 python tools/knn_stats.py : per-phase counters of the k-NN tile kernel on the config-2 catalogue."""
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from synthobs_b200 import synthetic
from synthobs_b200.morph import images
dev = torch.device('cuda', 0)
W, ndim = 512 * 0.031 + 1e-9, 1024
p = synthetic.particles(1000000, seed=200, scale_kpc=3.0)
X = torch.from_numpy(p['X'] - np.median(p['X'])).to(dev); Y = torch.from_numpy(p['Y'] - np.median(p['Y'])).to(dev)
width, _, G = images.pixel_grid_device(W / ndim, ndim)
ix, iy = images.ngp_index_device(X, Y, G, ndim, width / 2.)
for fl in (0.0, images.FWHM_FLOOR):
    for _ in range(2):
        images.knn_device(X, Y, ix, 8, width / 2., floor=fl)
    torch.cuda.synchronize()
