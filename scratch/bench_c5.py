"""BASELINE config 5: one huge object, N particles onto ndim x ndim, adaptive smoothing, particle shares
per rank + NCCL all-reduce.  torchrun --nproc-per-node G scratch/bench_c5.py [N] [ndim]"""
import os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as dist
from synthobs_b200 import dist as sdist
from synthobs_b200.morph import images
rank, ws = sdist.init_from_env('nccl')
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000000
ndim = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
K = 8
dev = torch.device('cuda', torch.cuda.current_device())
g = torch.Generator(device=dev); g.manual_seed(1234)
S = 10.0                    # kpc: the FWHM floor of 0.01 kpc (images.py:89) is then ~1 pixel
X = S * torch.randn(N, dtype=torch.float64, device=dev, generator=g)
Y = S * torch.randn(N, dtype=torch.float64, device=dev, generator=g)
m = N // 20
for c in range(3):          # three dense clumps of 5 % each
    cx, cy = (S * torch.randn(2, dtype=torch.float64, device=dev, generator=g)).tolist()
    sl = slice(c * m, (c + 1) * m)
    X[sl] = cx + 0.05 * S * torch.randn(m, dtype=torch.float64, device=dev, generator=g)
    Y[sl] = cy + 0.05 * S * torch.randn(m, dtype=torch.float64, device=dev, generator=g)
L = 8.47e5 * (0.5 + torch.rand(N, dtype=torch.float64, device=dev, generator=g))
res = 8.0 * S / ndim
def run():
    return sdist.image_huge(X, Y, L, res, ndim, K)
img = run()
torch.cuda.synchronize()
if ws > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 3
e0.record()
for _ in range(steps): img = run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
t = torch.tensor([ms], device=dev, dtype=torch.float64)
if ws > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
sel = ((X.abs() < 4.0 * S) & (Y.abs() < 4.0 * S))
flux = float(L[sel].sum()); tot = float(img.to(torch.float64).sum())
if rank == 0:
    print('C5 N=%.0e ndim=%d k=%d gpus=%d: %.1f ms/image -> %.3e particles/s ; flux conservation %.2e ; peak mem %.1f GB'
          % (N, ndim, K, ws, float(t), N / float(t) * 1e3, abs(tot / flux - 1), torch.cuda.max_memory_allocated() / 1e9))
    torch.save(img[::16, ::16].cpu(), 'gpurun_out/c5_img_%d.pt' % ws)
if ws > 1: dist.destroy_process_group()
