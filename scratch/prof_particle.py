"""host-side profile of images.particle (e2e path of the imaging bench)"""
import cProfile, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import bench
from synthobs_b200.morph import images

class A: pass
ctx = A(); ctx.rank = 0
p, filters, cosmo, width_arcsec, Lh, PSFs = bench.imaging_inputs(ctx)
Xh = torch.from_numpy(p['X']).pin_memory(); Yh = torch.from_numpy(p['Y']).pin_memory(); Lh_t = torch.from_numpy(Lh).pin_memory()
kw = dict(smoothing=('adaptive', 8.0), super_sampling=2)
Ld = {f: Lh_t[i] for i, f in enumerate(filters)}
def step():
    return images.particle(Xh, Yh, Ld, filters, cosmo, 8.0, width_arcsec, PSFs=PSFs, **kw)
for _ in range(3): step()
t0 = time.time()
for _ in range(10): step()
print('ms per call', (time.time() - t0) * 100)
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(25)
