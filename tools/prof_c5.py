"""config 5 per-rank cost on ONE GPU: the partial image of share r of N (what rank r of N would do, without the reduce)"""
import sys, time
import torch
sys.path.insert(0, '.')
from synthobs_b200 import _cabi, dist as sdist
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100000000
dev = torch.device('cuda', 0)
S, ndim = 10.0, 8192
g = torch.Generator(device=dev); g.manual_seed(99)
X = S * torch.randn(n, dtype=torch.float64, device=dev, generator=g)
Y = S * torch.randn(n, dtype=torch.float64, device=dev, generator=g)
m = n // 20
X[:m] = 3.0 + 0.5 * torch.randn(m, dtype=torch.float64, device=dev, generator=g)
Y[:m] = -7.0 + 0.5 * torch.randn(m, dtype=torch.float64, device=dev, generator=g)
L = 8.47e5 * (0.5 + torch.rand(n, dtype=torch.float64, device=dev, generator=g))
res = 8.0 * S / ndim
for r in range(N):
    for _ in range(2):
        sdist.image_huge(X, Y, L, res, ndim, 8, rank=r, n_ranks=N, reduce=False)
    torch.cuda.synchronize()
    st = {}
    t0 = time.time()
    for _ in range(3):
        sdist.image_huge(X, Y, L, res, ndim, 8, rank=r, n_ranks=N, reduce=False, stages=st)
    torch.cuda.synchronize()
    print('share %d/%d: %.2f ms' % (r, N, (time.time() - t0) * 1e3 / 3), {k[:28]: round(v / 3, 2) for k, v in st.items()})
_cabi.prof_enable(True)
sdist.image_huge(X, Y, L, res, ndim, 8, rank=N // 2, n_ranks=N, reduce=False)
torch.cuda.synchronize()
for k, (c, ms) in _cabi.prof_report().items():
    print('   %-36s %3d %8.3f ms' % (k, c, ms))
_cabi.prof_enable(False)
