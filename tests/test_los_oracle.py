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
