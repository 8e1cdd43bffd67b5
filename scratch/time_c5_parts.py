import sys; sys.path.insert(0, '.')
import numpy as np, torch
from synthobs_b200 import device as dv
from synthobs_b200.morph import images
dev = dv.device(); N = 100000000; ndim = 8192; S = 10.0
g = torch.Generator(device=dev); g.manual_seed(1234)
X = S * torch.randn(N, dtype=torch.float64, device=dev, generator=g); Y = S * torch.randn(N, dtype=torch.float64, device=dev, generator=g)
L = 8.47e5 * (0.5 + torch.rand(N, dtype=torch.float64, device=dev, generator=g)).reshape(1, -1)
res = 8.0 * S / ndim
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(3):
    t = [ev()]
    width, G = images.pixel_grid(res, ndim); Gd = dv.to_dev(G)
    ix, iy = images.ngp_index_device(X, Y, Gd, ndim, width / 2.); t.append(ev())
    dk, pids = images.knn_device_part(X, Y, ix, 8, width / 2., 0, 8, floor=images.FWHM_FLOOR); t.append(ev())
    Xs, Ys, Ls, dks = X[pids], Y[pids], L[:, pids], dk[pids]; t.append(ev())
    r = images.deposit_device(Xs, Ys, Ls, res, ndim, adaptive_k=8, want_ngp=False, dk=dks); t.append(ev())
    torch.cuda.synchronize()
print('share 1/8: ngp index (all) %.2f | knn tree build (all) + queries (1/8) %.2f | gathers %.2f | plan+deposit (1/8) %.2f ms' % tuple(t[i].elapsed_time(t[i+1]) for i in range(4)))
