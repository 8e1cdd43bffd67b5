import sys; sys.path.insert(0, '.')
import torch
from synthobs_b200 import los
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev); g.manual_seed(7)
ns = ng = 100000
stars = [1.5 * torch.randn(ns, dtype=torch.float64, device=dev, generator=g) for _ in range(3)]
gas = [2.0 * torch.randn(ng, dtype=torch.float64, device=dev, generator=g) for _ in range(3)]
m = 1e6 * (0.5 + torch.rand(ng, dtype=torch.float64, device=dev, generator=g))
Z = 10 ** (-4 + 2.3 * torch.rand(ng, dtype=torch.float64, device=dev, generator=g))
h = 10 ** (-1.3 + 1.3 * torch.rand(ng, dtype=torch.float64, device=dev, generator=g))
for _ in range(2):
    los.metal_surface_density(*stars, *gas, m, Z, h, scale=1e-6)
torch.cuda.synchronize()
