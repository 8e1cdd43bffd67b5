import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from synthobs_b200 import los
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev); g.manual_seed(7)
ns = ng = 100000
stars = [1.5 * torch.randn(ns, dtype=torch.float64, device=dev, generator=g) for _ in range(3)]
gas = [2.0 * torch.randn(ng, dtype=torch.float64, device=dev, generator=g) for _ in range(3)]
m = 1e6 * (0.5 + torch.rand(ng, dtype=torch.float64, device=dev, generator=g))
Z = 10 ** (-4 + 2.3 * torch.rand(ng, dtype=torch.float64, device=dev, generator=g))
h = 10 ** (-1.3 + 1.3 * torch.rand(ng, dtype=torch.float64, device=dev, generator=g))
import os
los.GRID_CELL = float(os.environ.get('SO_LOS_CELL', los.GRID_CELL))
print('GRID_CELL', los.GRID_CELL)
def T(f, n=5):
    for _ in range(2): r = f()
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(n): r = f()
    torch.cuda.synchronize(); return (time.time() - t0) * 1e3 / n, r
t, plan = T(lambda: los._cell_grid(stars[0], stars[1], gas[0], gas[1], h)); print('cell grid %.3f ms' % t)
order, so, gi, go = plan
print('objects', so.size - 1, 'entries', gi.numel(), 'max stars/cell', np.diff(so).max(), 'max gas/cell', np.diff(go).max())
t, gath = T(lambda: [a[gi] for a in gas + [m, Z, h]]); print('gather gas %.3f ms' % t)
t, sg = T(lambda: [a[order] for a in stars]); print('gather stars %.3f ms' % t)
t, res = T(lambda: los.metal_surface_density_batch(*sg, so, *gath, go, scale=1e-6)); print('batch kernel call %.3f ms' % t)
from synthobs_b200 import _cabi
_cabi.prof_enable(True); los.metal_surface_density_batch(*sg, so, *gath, go, scale=1e-6); torch.cuda.synchronize(); print(_cabi.prof_report()); _cabi.prof_enable(False)
