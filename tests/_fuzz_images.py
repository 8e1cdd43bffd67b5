"""Random small configurations: images.core / core_batch / k-NN vs the oracle.  Run on a B200."""
import sys
import numpy as np, torch
from oracle import image_oracle
from synthobs_b200 import device as dv, synthetic
from synthobs_b200.morph import images


def run(seed=0, cases=40):
    """returns the list of failing case numbers (details are printed)"""
    rng = np.random.default_rng(seed)
    bad = []
    for case in range(cases):
        n = int(rng.choice([1, 2, 3, 7, 33, 100, 257, 1000, 4000]))
        ndim = int(rng.choice([1, 2, 5, 16, 17, 33, 64, 100]))
        res = float(rng.choice([0.01, 0.1, 0.37, 2.0]))
        k = int(rng.choice([1, 2, 3, 8, 9, 16, 40]))
        mode = rng.choice(['ngp', 'adaptive', 'conv'])
        scale = ndim * res * float(rng.choice([0.05, 0.3, 1.0]))
        X = rng.normal(0, scale, n); Y = rng.normal(0, scale, n)
        if rng.random() < 0.3 and n > 4:        # duplicates and a tight clump
            X[: n // 2] = X[0] + rng.normal(0, 1e-9 * max(scale, 1e-3), n // 2); Y[: n // 2] = Y[0]
        if rng.random() < 0.2:
            X[rng.integers(n)] = np.inf if rng.random() < 0.5 else np.nan
        L = rng.lognormal(0, 1, n) * float(rng.choice([1e-20, 1.0, 1e25]))
        sm = False if mode == 'ngp' else (('adaptive', k) if mode == 'adaptive' else ('convolved_gaussian', float(res * rng.uniform(0.5, 4))))
        if mode == 'adaptive' and k == 1:
            continue       # reference raises IndexError for k = 1 (images.py:89)
        try:
            with np.errstate(all='ignore'):
                ref = image_oracle.core(X, Y, L, resolution=res, ndim=ndim, smoothing=sm)
                got = images.core(X.copy(), Y.copy(), L.copy(), resolution=res, ndim=ndim, smoothing=sm)
            ok = np.array_equal(got.hist, ref['hist'])
            peak = np.abs(ref['data']).max()
            if np.isfinite(peak) and peak > 0:
                ok = ok and np.abs(got.data - ref['data']).max() <= 2e-5 * peak
            if not ok:
                bad.append(case)
                print('MISMATCH case', case, dict(n=n, ndim=ndim, res=res, k=k, mode=str(mode), scale=scale),
                      'hist equal', np.array_equal(got.hist, ref['hist']), 'max diff/peak', np.abs(got.data - ref['data']).max() / peak if peak else None)
        except Exception as e:   # noqa: BLE001
            bad.append(case)
            print('ERROR case', case, dict(n=n, ndim=ndim, res=res, k=k, mode=str(mode)), repr(e)[:200])
    return bad



if __name__ == '__main__':
    failures = run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 60)
    print('fuzz done, failures:', len(failures))
