"""TEST INFRASTRUCTURE — CPU (NumPy, float64) SPECIFICATION and oracle of the line-of-sight metal column density
(SURVEY.md §8 f4).  Not shipped, not a fallback: only tests/ and bench.py's cpu_baseline leg import this module.

Parity status: **parity unpinned against the reference** — the reference does not contain this computation at all:
`MetSurfaceDensities.npy` is an INPUT of its loaders (/root/reference/synthobs/core.py:17-21, :33-37) that the examples
turn into the ISM optical depth with `tauVs_ISM = (10**5.2) * MetSurfaceDensities` (examples/SED_luminosities.py:69,
examples/SED_fluxes.py).  This file IS the specification the CUDA path (synthobs_b200/csrc/los.cu) is checked against;
the kernel table below is pinned by its own known answers (normalisation, value at the centre, direct quadrature).

Specification.  For every star particle i at (x_i, y_i, z_i), looking along +z,

    Sigma_i = scale * sum over gas particles j with z_j > z_i of  m_j Z_j / h_j^2 * F(b_ij / h_j),
    b_ij = sqrt((x_i - x_j)^2 + (y_i - y_j)^2),

i.e. the metal mass of the gas IN FRONT of the star, smoothed with each gas particle's SPH kernel projected along the
line of sight.  W(r, h) = w(r / h) / h^3 is the cubic spline with compact support r < h (the Gadget convention):

    w(q) = 8 / pi * (1 - 6 q^2 + 6 q^3)      0   <= q <= 1/2
    w(q) = 8 / pi * 2 (1 - q)^3              1/2 <  q <= 1
    w(q) = 0                                 q > 1

and F(u) = 2 * integral_0^sqrt(1 - u^2) w(sqrt(u^2 + t^2)) dt is its projection (integral of F(u) 2 pi u du over
0..1 = 1).  F is TABULATED: F_tab[i] = F(i / NT), i = 0..NT, NT = 1024, computed in float64 by Gauss-Legendre
quadrature; between nodes F is interpolated linearly, and F(u) = 0 for u >= 1.  The z comparison is strict and made on
the float64 inputs.  `scale` converts units (1 for M_sol / kpc^2 with kpc and M_sol inputs; 1e-6 for M_sol / pc^2).
A whole particle counts or does not (no partial kernel overlap along z): the same simplification the simulation
pipelines that wrote MetSurfaceDensities make.
"""

import numpy as np

NT = 1024


def w3(q):
    """dimensionless cubic spline, support q < 1 (Gadget convention), integral of w over the unit ball = 1"""
    q = np.asarray(q, dtype=np.float64)
    out = np.zeros_like(q)
    a = q <= 0.5
    b = (q > 0.5) & (q <= 1.0)
    out[a] = 1.0 - 6.0 * q[a] ** 2 + 6.0 * q[a] ** 3
    out[b] = 2.0 * (1.0 - q[b]) ** 3
    return out * (8.0 / np.pi)


def projected_kernel(u, order=96):
    """F(u) = 2 int_0^sqrt(1-u^2) w(sqrt(u^2 + t^2)) dt by Gauss-Legendre quadrature, split at the kink q = 1/2"""
    u = np.atleast_1d(np.asarray(u, dtype=np.float64))
    x, wq = np.polynomial.legendre.leggauss(order)
    out = np.zeros_like(u)
    for n, un in enumerate(u):
        if un >= 1.0:
            continue
        tmax = np.sqrt(1.0 - un * un)
        tmid = np.sqrt(0.25 - un * un) if un < 0.5 else 0.0
        total = 0.0
        for lo, hi in ((0.0, tmid), (tmid, tmax)):
            if hi <= lo:
                continue
            t = 0.5 * (hi - lo) * x + 0.5 * (hi + lo)
            total += 0.5 * (hi - lo) * np.sum(wq * w3(np.sqrt(un * un + t * t)))
        out[n] = 2.0 * total
    return out


_table = {}


def kernel_table(nt=NT):
    """F_tab[i] = F(i / nt), i = 0..nt (float64; F_tab[nt] = 0)"""
    if nt not in _table:
        t = projected_kernel(np.arange(nt + 1) / float(nt))
        t[nt] = 0.0
        _table[nt] = t
    return _table[nt]


def interp_kernel(u, table=None):
    """the specification's F: linear interpolation of the table, 0 for u >= 1"""
    table = kernel_table() if table is None else table
    nt = table.size - 1
    u = np.asarray(u, dtype=np.float64)
    t = np.clip(u, 0.0, 1.0) * nt
    i = np.minimum(t.astype(np.int64), nt - 1)
    f = t - i
    out = table[i] + f * (table[i + 1] - table[i])
    return np.where(u < 1.0, out, 0.0)


def metal_surface_density(xs, ys, zs, xg, yg, zg, mg, Zg, hg, scale=1.0, chunk=256, exact_kernel=False):
    """Sigma_i of the specification above for one object.  exact_kernel=True evaluates F by quadrature instead of the
    table (tests bound the interpolation error with it)."""
    xs, ys, zs = (np.asarray(a, dtype=np.float64) for a in (xs, ys, zs))
    xg, yg, zg, mg, Zg, hg = (np.asarray(a, dtype=np.float64) for a in (xg, yg, zg, mg, Zg, hg))
    out = np.zeros(xs.size)
    if xg.size == 0:
        return out
    wh2 = mg * Zg / (hg * hg)
    for a in range(0, xs.size, chunk):
        b = min(xs.size, a + chunk)
        dx = xs[a:b, None] - xg[None, :]
        dy = ys[a:b, None] - yg[None, :]
        u = np.sqrt(dx * dx + dy * dy) / hg[None, :]
        front = zg[None, :] > zs[a:b, None]
        F = projected_kernel(u.ravel()).reshape(u.shape) if exact_kernel else interp_kernel(u)
        out[a:b] = np.sum(np.where(front, wh2[None, :] * F, 0.0), axis=1)
    return scale * out


def metal_surface_density_batch(stars, star_off, gas, gas_off, scale=1.0):
    """loop of metal_surface_density over the objects of a CSR batch: stars = (xs, ys, zs), gas = (xg, yg, zg, mg, Zg, hg)"""
    out = np.zeros(np.asarray(stars[0]).size)
    for g in range(len(star_off) - 1):
        s0, s1, g0, g1 = int(star_off[g]), int(star_off[g + 1]), int(gas_off[g]), int(gas_off[g + 1])
        out[s0:s1] = metal_surface_density(*(np.asarray(a)[s0:s1] for a in stars), *(np.asarray(a)[g0:g1] for a in gas),
                                           scale=scale)
    return out
