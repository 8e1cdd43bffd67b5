import sys; sys.path.insert(0, '.')
import numpy as np, torch
from synthobs_b200.morph import images
from synthobs_b200 import synthetic
p = synthetic.particles(1000000, seed=1)
X, Y = torch.from_numpy(p['X']).cuda(), torch.from_numpy(p['Y']).cuda()
XY = torch.stack([X, Y]); X, Y = XY[0], XY[1]
for _ in range(5): images.median_device_pair(X, Y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): mx, my = images.median_device_pair(X, Y)
e1.record(); torch.cuda.synchronize()
print('median pair 1e6: %.4f ms' % (e0.elapsed_time(e1) / 50), float(mx) == np.median(p['X']), float(my) == np.median(p['Y']))
Z = torch.zeros(1000000, dtype=torch.float64, device='cuda')
for _ in range(2): images.median_device(Z)
torch.cuda.synchronize(); e0.record()
for _ in range(5): images.median_device(Z)
e1.record(); torch.cuda.synchronize()
print('degenerate (all equal) 1e6: %.4f ms' % (e0.elapsed_time(e1) / 5))
