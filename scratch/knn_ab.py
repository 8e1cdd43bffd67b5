"""k-NN A/B aid: time so_knn_kth_distance_ex on the bench catalogue (config 2) and a uniform one"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from synthobs_b200 import synthetic
from synthobs_b200.morph import images
from synthobs_b200 import device as dv
k = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for name, p in (('bench', synthetic.particles(1000000, seed=200, scale_kpc=3.0)),):
    X = torch.from_numpy(p['X'] - np.median(p['X'])).cuda(); Y = torch.from_numpy(p['Y'] - np.median(p['Y'])).cuda()
    width, G, Gd = images.pixel_grid_device(0.0762, 1024)
    ix, iy = images.ngp_index_device(X, Y, Gd, 1024, width / 2.)
    for floor in (0.0, 0.01):
        for _ in range(3): dk = images.knn_device(X, Y, ix, k, width / 2., floor=floor)
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): dk = images.knn_device(X, Y, ix, k, width / 2., floor=floor)
        e1.record(); torch.cuda.synchronize()
        print(name, 'k', k, 'floor', floor, 'ms', e0.elapsed_time(e1) / 10)
