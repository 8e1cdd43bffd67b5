import sys; sys.path.insert(0, '.')
import numpy as np, torch
from synthobs_b200 import _cabi
from synthobs_b200.sed import models
# rounding direction: all-negative A
rng = np.random.default_rng(1)
A = -rng.uniform(0.5, 1.5, (256, 64)); B = rng.uniform(0.5, 1.5, (256, 64))
Ad = torch.from_numpy(A.astype(np.float32)).cuda(); Bd = torch.from_numpy(B.astype(np.float32)).cuda()
ref = Ad.double().cpu().numpy() @ Bd.double().cpu().numpy().T
C = models.gemm(Ad, Bd).double().cpu().numpy()
print('negative A: mean (C-ref)/|ref| = %+.2e  (positive -> truncation toward zero; negative -> floor)' % ((C - ref) / np.abs(ref)).mean())
for (M, N, K) in ((2, 9000, 1326), (2048, 9000, 1326), (20000, 9000, 1326), (8192, 8192, 8192)):
    A = torch.rand((M, (K + 3) // 4 * 4), device='cuda'); B = torch.rand((N, (K + 3) // 4 * 4), device='cuda')
    Ah, Al = models.split_tf32(A); Bh, Bl = models.split_tf32(B)
    C = torch.empty((M, N), device='cuda')
    def run():
        _cabi.check(_cabi.lib.so_gemm_3xtf32(_cabi.ptr(Ah), _cabi.ptr(Al), A.stride(0), _cabi.ptr(Bh), _cabi.ptr(Bl), B.stride(0), _cabi.ptr(C), C.stride(0), M, N, K, _cabi.stream()), 'gemm')
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = 20 if M < 8000 else 5
    e0.record()
    for _ in range(it): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / it
    print('M%6d N%5d K%5d: %.3f ms  algorithmic %.1f TFLOP/s  (executed tf32 x3: %.1f TFLOP/s)' % (M, N, K, ms, 2 * M * N * K / ms / 1e9, 6 * M * N * K / ms / 1e9))
