import sys; sys.path.insert(0, '.')
import numpy as np, torch
from synthobs_b200.sed import models
def run(M, N, K, kind):
    rng = np.random.default_rng(1)
    if kind == 'pos1':
        A = rng.uniform(0.5, 1.5, (M, K)); B = rng.uniform(0.5, 1.5, (N, K))
    elif kind == 'wide':
        A = 10.0 ** rng.uniform(-3, 3, (M, K)); B = 10.0 ** rng.uniform(-2, 2, (N, K))
    elif kind == 'signed':
        A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K))
    Ad = torch.from_numpy(A.astype(np.float32)).cuda(); Bd = torch.from_numpy(B.astype(np.float32)).cuda()
    ref = Ad.double().cpu().numpy() @ Bd.double().cpu().numpy().T
    scale = np.abs(Ad.double().cpu().numpy()) @ np.abs(Bd.double().cpu().numpy()).T
    out = {}
    for algo in ('3xtf32', 'f32'):
        C = models.gemm(Ad, Bd, algo=algo).double().cpu().numpy()
        e = (C - ref) / scale
        out[algo] = (e.mean(), np.abs(e).max())
    # hi only
    Ah, Al = models.split_tf32(Ad); Bh, Bl = models.split_tf32(Bd)
    refh = Ah.double().cpu().numpy() @ Bh.double().cpu().numpy().T
    print('%-7s M%5d N%5d K%5d | 3xtf32 mean %+.2e max %.2e | f32 mean %+.2e max %.2e | hi*hi-only model err %.2e' % (
        kind, M, N, K, *out['3xtf32'], *out['f32'], np.abs((refh - ref) / scale).max()))
for kind in ('pos1', 'wide', 'signed'):
    for K in (32, 64, 128, 1024, 9000):
        run(256, 256, K, kind)
