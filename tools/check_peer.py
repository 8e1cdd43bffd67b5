"""multi-GPU check (torchrun, >= 2 ranks): dist.image_huge with the peer-memory block exchange == with ncclAllReduce
== the single-rank image; prints the timings of both reduce paths"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
from synthobs_b200 import dist as sdist
from synthobs_b200.morph import images
rank, ws = sdist.init_from_env()
dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
n, ndim, S = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3000000, int(sys.argv[2]) if len(sys.argv) > 2 else 4096, 10.0
g = torch.Generator(device=dev); g.manual_seed(5)
X = S * torch.randn(n, dtype=torch.float64, device=dev, generator=g)
Y = S * torch.randn(n, dtype=torch.float64, device=dev, generator=g)
X[:n // 20] = 3.0 + 0.3 * torch.randn(n // 20, dtype=torch.float64, device=dev, generator=g)
L = torch.stack([0.5 + torch.rand(n, dtype=torch.float64, device=dev, generator=g) for _ in range(2)])
res = 8.0 * S / ndim
out = {}
for mode in ('1', '0'):
    os.environ['SO_PEER_REDUCE'] = mode
    for _ in range(2):
        img = sdist.image_huge(X, Y, L, res, ndim, 8)
    torch.cuda.synchronize(); dist.barrier(); t0 = time.time()
    for _ in range(5):
        img = sdist.image_huge(X, Y, L, res, ndim, 8)
    torch.cuda.synchronize(); dist.barrier()
    out[mode] = (img.clone(), (time.time() - t0) * 200)
from synthobs_b200 import _cabi
os.environ['SO_PEER_REDUCE'] = '1'
_cabi.prof_enable(True)
st = {}
sdist.image_huge(X, Y, L, res, ndim, 8, stages=st)
torch.cuda.synchronize()
if rank == 0:
    print({k[:40]: round(v, 3) for k, v in st.items()})
    print({k: round(ms, 3) for k, (c, ms) in _cabi.prof_report().items() if 'peer' in k or 'mask' in k or 'push' in k or 'deposit' in k or 'tile_reduce' in k})
_cabi.prof_enable(False)
whole = sdist.image_huge(X, Y, L, res, ndim, 8, rank=0, n_ranks=1, reduce=False)
peak = float(whole.max())
d_peer = float((out['1'][0] - whole).abs().max()) / peak
d_nccl = float((out['0'][0] - whole).abs().max()) / peak
same = [torch.empty_like(out['1'][0]) for _ in range(ws)]
dist.all_gather(same, out['1'][0])
ident = all(torch.equal(same[0], t) for t in same)
if rank == 0:
    print('ranks %d  n %g  ndim %d: peer %.2f ms, nccl %.2f ms per call; |peer - 1 rank| / peak = %.2e, |nccl - 1 rank| / peak = %.2e, '
          'identical on all ranks: %s' % (ws, n, ndim, out['1'][1], out['0'][1], d_peer, d_nccl, ident))
    assert d_peer < 2e-6 and d_nccl < 2e-6 and ident
dist.destroy_process_group()
