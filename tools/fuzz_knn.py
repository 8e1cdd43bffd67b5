"""Random k-NN cases answered by the search selected with SO_KNN_ALGO (read once per process): run once with
SO_KNN_ALGO=quadtree and once with SO_KNN_ALGO=tiles, then `python tools/fuzz_knn.py --compare a.npz b.npz` checks that the
exact distances (floor = 0) are bit-identical and that the floored ones agree above the floor.

    SO_KNN_ALGO=quadtree python tools/fuzz_knn.py out_q.npz ; SO_KNN_ALGO=tiles python tools/fuzz_knn.py out_t.npz
"""
import sys
sys.path.insert(0, '.')
import numpy as np

if sys.argv[1] == '--compare':
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    bad = 0
    for key in a.files:
        x, y = a[key], b[key]
        if key.startswith('exact'):
            ok = np.array_equal(x, y)
        else:
            fl = 0.01
            ok = np.array_equal(np.maximum(x, fl), np.maximum(y, fl))
        if not ok:
            bad += 1
            d = np.flatnonzero(~((x == y) | (np.isinf(x) & np.isinf(y))))
            print('MISMATCH', key, d.size, x[d[:3]], y[d[:3]])
    print('%d arrays compared, %d mismatches' % (len(a.files), bad))
    sys.exit(1 if bad else 0)

import torch
from synthobs_b200 import device as dv
from synthobs_b200.morph import images
rng = np.random.default_rng(12345)
out = {}
for case in range(160):
    n = int(rng.choice([1, 2, 7, 31, 32, 33, 100, 1000, 5000, 20000, 32768, 40000]))
    n = max(1, int(n * rng.uniform(0.6, 1.0))) if n > 40 else n
    k = int(rng.choice([1, 2, 3, 5, 8, 9, 16, 17, 32]))
    kind = case % 5
    if kind == 0:
        P = rng.normal(0, 1.0, (n, 2))
    elif kind == 1:                                   # clustered + duplicates
        P = rng.normal(0, 1.5, (n, 2)); m = n // 3
        P[:m] = rng.normal(0.7, 0.01, (m, 2)); P[m:m + m // 4] = P[:m // 4]
    elif kind == 2:                                   # lattice: exact ties
        s = int(np.ceil(np.sqrt(n))); g = np.stack(np.meshgrid(np.arange(s), np.arange(s)), -1).reshape(-1, 2)[:n]
        P = g * 0.03125 - 1.0
    elif kind == 3:                                   # streak + outliers (some outside the image)
        P = np.column_stack([np.linspace(-3.5, 3.5, n), 1e-7 * rng.standard_normal(n)]); P[::17] = rng.uniform(-6, 6, (len(P[::17]), 2))
    else:                                             # many images in one batch
        P = rng.normal(0, 0.8, (n, 2))
    X, Y = np.ascontiguousarray(P[:, 0]), np.ascontiguousarray(P[:, 1])
    hw = 4.0
    width, G = images.pixel_grid(2 * hw / 100, 100)
    Xd, Yd = dv.to_dev(X), dv.to_dev(Y)
    ix, _ = images.ngp_index_device(Xd, Yd, dv.to_dev(G), 100, hw)
    img, n_img = None, 1
    if kind == 4 and n >= 8:
        n_img = int(rng.integers(2, min(n, 400)))
        ids = np.sort(rng.integers(0, n_img, n)).astype(np.int32)
        img = dv.to_dev(ids, torch.int32)
    if '-v' in sys.argv:
        print(case, n, k, kind, n_img, flush=True)
    out['exact_%03d_n%d_k%d_t%d' % (case, n, k, kind)] = images.knn_device(Xd, Yd, ix, k, hw, img=img, n_img=n_img).cpu().numpy()
    out['floor_%03d_n%d_k%d_t%d' % (case, n, k, kind)] = images.knn_device(Xd, Yd, ix, k, hw, img=img, n_img=n_img, floor=0.01).cpu().numpy()
    torch.cuda.synchronize()
np.savez(sys.argv[1], **out)
print('wrote', len(out), 'arrays')
