"""CUDA line-of-sight metal column density (so_los_column) against the specification oracle/los_oracle.py.
Tolerance: the pair arithmetic is FP32 (relative coordinates, kernel interpolation), summed per tile of 128 gas particles
in FP32 and over the tiles in float64: |a - b| <= 2e-5 max|b| and 1e-5 in L2."""
import numpy as np
import pytest

from oracle import los_oracle as lo

pytestmark = pytest.mark.gpu


def _close(a, b):
    assert np.max(np.abs(a - b)) <= 2e-5 * np.max(np.abs(b)) + 1e-300
    assert np.linalg.norm(a - b) <= 1e-5 * np.linalg.norm(b) + 1e-300


def _galaxy(rng, ns, ng, centre=(0.0, 0.0, 0.0)):
    stars = rng.normal(0, 1.5, (3, ns)) + np.asarray(centre)[:, None]
    gas = rng.normal(0, 2.0, (3, ng)) + np.asarray(centre)[:, None]
    m = rng.uniform(0.5e6, 2e6, ng)
    Z = 10 ** rng.uniform(-4, -1.7, ng)
    h = 10 ** rng.uniform(-1.3, 0.0, ng)
    return stars, gas, m, Z, h


@pytest.mark.parametrize('ns,ng', [(1, 1), (300, 1000), (5000, 20000), (257, 255)])
def test_los_single_object_matches_the_specification(ns, ng):
    from synthobs_b200 import los
    rng = np.random.default_rng(ns + ng)
    stars, gas, m, Z, h = _galaxy(rng, ns, ng, centre=(40.0, -25.0, 3.0))     # far from the origin: relative coordinates matter
    out = los.metal_surface_density(*stars, *gas, m, Z, h, scale=1e-6)
    ref = lo.metal_surface_density(*stars, *gas, m, Z, h, scale=1e-6)
    assert out.dtype == np.float64 and out.shape == (ns,)
    _close(out, ref)


def test_los_edges_and_properties():
    import torch
    from synthobs_b200 import los
    rng = np.random.default_rng(11)
    stars, gas, m, Z, h = _galaxy(rng, 400, 3000)
    # a star exactly at a gas particle's position and depth: that particle is not in front (strict), one just behind is
    stars[:, 0] = gas[:, 0]
    stars[:, 1] = gas[:, 1]
    stars[2, 1] = np.nextafter(gas[2, 1], -np.inf)
    ref = lo.metal_surface_density(*stars, *gas, m, Z, h)
    out = los.metal_surface_density(*stars, *gas, m, Z, h)
    _close(out, ref)
    # no gas at all / no gas in front
    assert np.array_equal(los.metal_surface_density(*stars, *(np.zeros((3, 0))), np.zeros(0), np.zeros(0), np.zeros(0)), np.zeros(400))
    behind = gas.copy()
    behind[2] = stars[2].min() - 1.0
    assert np.array_equal(los.metal_surface_density(*stars, *behind, m, Z, h), np.zeros(400))
    # linear in the metal mass; bitwise reproducible; CUDA tensors in -> CUDA tensor out
    twice = los.metal_surface_density(*stars, *gas, 2.0 * m, Z, h)
    assert np.array_equal(twice, 2.0 * out)
    assert np.array_equal(los.metal_surface_density(*stars, *gas, m, Z, h), out)
    dev = torch.device('cuda', 0)
    t = los.metal_surface_density(*(torch.from_numpy(a).to(dev) for a in stars), *(torch.from_numpy(a).to(dev) for a in gas),
                                  torch.from_numpy(m).to(dev), torch.from_numpy(Z).to(dev), torch.from_numpy(h).to(dev))
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), out)


def test_los_batch_is_a_loop_over_objects():
    from synthobs_b200 import los
    rng = np.random.default_rng(5)
    sizes = [(0, 50), (700, 0), (1, 3000), (513, 800), (2000, 5000), (30, 30)]
    S, G, M, ZZ, H = [], [], [], [], []
    for k, (ns, ng) in enumerate(sizes):
        stars, gas, m, Z, h = _galaxy(rng, ns, ng, centre=(10.0 * k, -3.0 * k, 0.5 * k))
        S.append(stars); G.append(gas); M.append(m); ZZ.append(Z); H.append(h)
    stars, gas = np.concatenate(S, axis=1), np.concatenate(G, axis=1)
    m, Z, h = np.concatenate(M), np.concatenate(ZZ), np.concatenate(H)
    so = np.concatenate([[0], np.cumsum([s[0] for s in sizes])])
    go = np.concatenate([[0], np.cumsum([s[1] for s in sizes])])
    out = los.metal_surface_density_batch(*stars, so, *gas, m, Z, h, go, scale=1e-6)
    ref = lo.metal_surface_density_batch(tuple(stars), so, (*gas, m, Z, h), go, scale=1e-6)
    _close(out, ref)
    for k in range(len(sizes)):      # the single-object call on the same particles (the gas chunks of a call depend on its size: float64 sum order)
        one = los.metal_surface_density(*stars[:, so[k]:so[k + 1]], *gas[:, go[k]:go[k + 1]], m[go[k]:go[k + 1]], Z[go[k]:go[k + 1]],
                                        h[go[k]:go[k + 1]], scale=1e-6)
        np.testing.assert_allclose(one, out[so[k]:so[k + 1]], rtol=1e-13, atol=0)
    with pytest.raises(ValueError):
        los.metal_surface_density_batch(*stars, so[:-1], *gas, m, Z, h, go, scale=1.0)


def test_los_cell_grid_equals_all_pairs():
    """grid=True bins the gas per cell of the stars' grid and runs every cell as one object of the same kernel: same
    numbers as all pairs (float64 sum order over tiles differs) and as the specification"""
    from synthobs_b200 import los
    rng = np.random.default_rng(21)
    stars, gas, m, Z, h = _galaxy(rng, 20000, 30000, centre=(40.0, -25.0, 3.0))
    a = los.metal_surface_density(*stars, *gas, m, Z, h, scale=1e-6, grid=True)
    b = los.metal_surface_density(*stars, *gas, m, Z, h, scale=1e-6, grid=False)
    np.testing.assert_allclose(a, b, rtol=2e-6, atol=1e-9 * b.max())    # FP32 tile sums are cut differently
    ref = lo.metal_surface_density(*stars[:, :2000], *gas, m, Z, h, scale=1e-6)
    _close(a[:2000], ref)
    # one star, and gas far larger than the stars' box
    one = los.metal_surface_density(*stars[:, :1], *gas, m, Z, np.full(h.size, 3.0), grid=True)
    _close(one, lo.metal_surface_density(*stars[:, :1], *gas, m, Z, np.full(h.size, 3.0)))
