"""Line-of-sight metal column density: the specification in oracle/los_oracle.py is pinned by its
own known answers (the reference holds no implementation: MetSurfaceDensities is an input of synthobs/core.py:17-21)."""
import ast
import os

import numpy as np

from oracle import los_oracle as lo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_projected_kernel_known_answers():
    t = lo.kernel_table()
    # F(0) = 2 int_0^1 w(t) dt = 16/pi (11/32 + 1/32) = 6/pi
    assert abs(t[0] - 6.0 / np.pi) < 1e-14
    assert t[-1] == 0.0 and np.all(np.diff(t) <= 0.0) and np.all(t >= 0.0)
    # the projection conserves the kernel's unit volume: int F(u) 2 pi u du = 1
    u = np.linspace(0.0, 1.0, 400001)
    assert abs(np.trapezoid(lo.projected_kernel(u[::400]) * 2 * np.pi * u[::400], u[::400]) - 1.0) < 1e-5
    assert abs(np.trapezoid(lo.interp_kernel(u) * 2 * np.pi * u, u) - 1.0) < 2e-6
    # 3-D normalisation of w itself
    q = np.linspace(0.0, 1.0, 200001)
    assert abs(np.trapezoid(lo.w3(q) * 4 * np.pi * q * q, q) - 1.0) < 1e-9
    # linear interpolation of 1025 nodes: error below 2e-6 of the peak
    um = (np.arange(lo.NT) + 0.5) / lo.NT
    assert np.max(np.abs(lo.interp_kernel(um) - lo.projected_kernel(um))) < 2e-6 * t[0]


def test_oracle_geometry():
    # one gas particle in front, on the axis: Sigma = m Z / h^2 * 6 / pi; behind: nothing; outside the support: nothing
    one = lo.metal_surface_density([0.0], [0.0], [0.0], [0.0], [0.0], [1.0], [3.0], [0.01], [0.5])
    assert abs(one[0] - 3.0 * 0.01 / 0.25 * 6.0 / np.pi) < 1e-15
    assert lo.metal_surface_density([0.0], [0.0], [0.0], [0.0], [0.0], [-1.0], [3.0], [0.01], [0.5])[0] == 0.0
    assert lo.metal_surface_density([0.0], [0.0], [0.0], [0.0], [0.0], [0.0], [3.0], [0.01], [0.5])[0] == 0.0   # z_j > z_i is strict
    assert lo.metal_surface_density([0.0], [0.0], [0.0], [0.5], [0.0], [1.0], [3.0], [0.01], [0.5])[0] == 0.0
    # a uniform slab of gas: the column of a star below it is the slab's metal surface density
    rng = np.random.default_rng(3)
    n, L = 40000, 4.0
    xg, yg = rng.uniform(-L, L, (2, n))
    zg = rng.uniform(1.0, 2.0, n)
    sig = lo.metal_surface_density([0.0], [0.0], [0.0], xg, yg, zg, np.full(n, 2.0), np.full(n, 0.02), np.full(n, 0.8))
    assert abs(sig[0] / (n * 2.0 * 0.02 / (2 * L) ** 2) - 1.0) < 0.05
    # table vs direct quadrature on a random configuration; batch = loop; scale is a factor
    xs, ys, zs = rng.normal(0, 1, (3, 60))
    g = rng.normal(0, 1, (3, 500))
    m, Z, h = rng.uniform(1, 2, 500), rng.uniform(0, 0.02, 500), rng.uniform(0.1, 0.6, 500)
    a = lo.metal_surface_density(xs, ys, zs, *g, m, Z, h)
    b = lo.metal_surface_density(xs, ys, zs, *g, m, Z, h, exact_kernel=True)
    assert np.max(np.abs(a - b)) < 2e-6 * a.max()
    c = lo.metal_surface_density_batch((xs, ys, zs), [0, 25, 25, 60], (*g, m, Z, h), [0, 200, 300, 500], scale=1e-6)
    assert np.allclose(c[:25], 1e-6 * lo.metal_surface_density(xs[:25], ys[:25], zs[:25], *(v[:200] for v in g), m[:200], Z[:200], h[:200]),
                       rtol=1e-14, atol=0)
    assert np.allclose(c[25:], 1e-6 * lo.metal_surface_density(xs[25:], ys[25:], zs[25:], *(v[300:] for v in g), m[300:], Z[300:], h[300:]),
                       rtol=1e-14, atol=0)


def test_product_table_is_the_specified_one():
    """synthobs_b200/los.py builds its own table (the product never imports oracle/): the same numbers"""
    src = open(os.path.join(ROOT, 'synthobs_b200', 'los.py')).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == '_projected_spline'][0]
    ns = {'np': np}
    exec(compile(ast.Module([fn], []), 'los.py', 'exec'), ns)
    u = np.arange(lo.NT + 1) / float(lo.NT)
    assert np.max(np.abs(ns['_projected_spline'](u) - lo.projected_kernel(u))) < 1e-14


def test_cell_grid_is_exact():
    """the cell grid of synthobs_b200/los.py (torch ops; run here on CPU tensors): every star's cell list holds all the gas
    whose kernel can cover it, so the oracle over the cells equals the oracle over all pairs (float64 sum order only)"""
    import torch
    src = open(os.path.join(ROOT, 'synthobs_b200', 'los.py')).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == '_cell_grid'][0]
    ns_ = {'np': np, 'torch': torch, 'GRID_MAX_ENTRIES': 4.0e7, 'GRID_CELL': 4.0}
    exec(compile(ast.Module([fn], []), 'los.py', 'exec'), ns_)
    rng = np.random.default_rng(1)
    for ns, ng, hlo in ((3000, 8000, -1.3), (500, 4000, -2.0), (1, 300, -0.5)):
        xs, ys, zs = rng.normal(0, 1.5, (3, ns)) + np.array([[40.0], [-25.0], [3.0]])
        xg, yg, zg = rng.normal(0, 2.0, (3, ng)) + np.array([[40.0], [-25.0], [3.0]])
        m, Z, h = rng.uniform(0.5e6, 2e6, ng), 10 ** rng.uniform(-4, -1.7, ng), 10 ** rng.uniform(hlo, 0.0, ng)
        T = torch.from_numpy
        order, so, gi, go = ns_['_cell_grid'](T(xs), T(ys), T(xg), T(yg), T(h))
        order, gi = order.numpy(), gi.numpy()
        assert so[-1] == ns and go[-1] == gi.size and np.array_equal(np.sort(order), np.arange(ns))
        o = lo.metal_surface_density_batch((xs[order], ys[order], zs[order]), so, (xg[gi], yg[gi], zg[gi], m[gi], Z[gi], h[gi]), go)
        out = np.empty(ns)
        out[order] = o
        ref = lo.metal_surface_density(xs, ys, zs, xg, yg, zg, m, Z, h)
        assert np.max(np.abs(out - ref)) <= 1e-13 * max(ref.max(), 1e-300)
        if ns > 1:
            assert np.sum(np.diff(so) * np.diff(go)) < 0.7 * ns * ng          # and it saves work
    # entry budget exceeded -> None (the caller takes all pairs)
    ns_['GRID_MAX_ENTRIES'] = 10.0
    assert ns_['_cell_grid'](T(xs), T(ys), T(xg), T(yg), T(np.full(ng, 50.0))) is None
