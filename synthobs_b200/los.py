"""Line-of-sight metal column density of star particles: the quantity the reference's loaders read from
`MetSurfaceDensities.npy` (reference: synthobs/core.py:17-21, :33-37) and its examples turn into the ISM optical depth,
`tauVs_ISM = (10**5.2) * MetSurfaceDensities` (examples/SED_luminosities.py:69).  The reference itself never computes it
(SURVEY.md section 8 f4); the specification this module implements is written out in oracle/los_oracle.py:

    Sigma_i = scale * sum_{gas j in front of star i: z_j > z_i}  m_j Z_j / h_j^2 * F(b_ij / h_j)

with F the cubic-spline SPH kernel (support r < h) projected along the line of sight (+z), tabulated at 1025 nodes and
interpolated linearly.  CUDA only (so_los_column, csrc/los.cu): there is no CPU path.

    from synthobs_b200 import los
    Sigma = los.metal_surface_density(xs, ys, zs, xg, yg, zg, gas_mass, gas_Z, gas_h, scale=1e-6)   # M_sol / pc^2
    tauVs_ISM = 10 ** 5.2 * Sigma
"""

import ctypes

import numpy as np
import torch

from . import _cabi
from . import device as dv

NT = 1024
_tables = {}


def _projected_spline(u, order=96):
    """F(u) = 2 int_0^sqrt(1 - u^2) w(sqrt(u^2 + t^2)) dt, w(q) = 8/pi (1 - 6 q^2 + 6 q^3 | 2 (1 - q)^3 | 0) for
    q <= 1/2 | <= 1 | beyond: Gauss-Legendre on the two smooth pieces of the path (split where q = 1/2), float64."""
    u = np.asarray(u, dtype=np.float64)
    x, wq = np.polynomial.legendre.leggauss(order)
    inside = u < 1.0
    uu = np.where(inside, u, 0.0)
    t_end = np.sqrt(np.clip(1.0 - uu * uu, 0.0, None))
    t_mid = np.sqrt(np.clip(0.25 - uu * uu, 0.0, None))
    total = np.zeros_like(uu)
    for a, b in ((np.zeros_like(uu), t_mid), (t_mid, t_end)):
        half = 0.5 * (b - a)
        t = half[:, None] * x[None, :] + 0.5 * (a + b)[:, None]
        q = np.sqrt(uu[:, None] ** 2 + t * t)
        w = np.where(q <= 0.5, 1.0 - 6.0 * q * q + 6.0 * q ** 3, 2.0 * np.clip(1.0 - q, 0.0, None) ** 3)
        total += half * (w * wq[None, :]).sum(axis=1)
    return np.where(inside, 2.0 * (8.0 / np.pi) * total, 0.0)


def kernel_table(nt=NT):
    """(float64 host table F(i / nt), i = 0..nt; the same as a float32 CUDA tensor), cached per device"""
    dev = dv.device()
    key = (int(nt), dev.index)
    if key not in _tables:
        t = _projected_spline(np.arange(nt + 1) / float(nt))
        t[nt] = 0.0
        _tables[key] = (t, torch.from_numpy(t.astype(np.float32)).to(dev))
    return _tables[key]


def metal_surface_density_batch(xs, ys, zs, star_offsets, xg, yg, zg, gas_mass, gas_Z, gas_h, gas_offsets, scale=1.0):
    """Sigma for a batch of objects: stars and gas concatenated object after object, `star_offsets` / `gas_offsets` =
    CSR offsets [n_obj + 1] (a star only sees the gas of its own object).  Inputs: NumPy arrays or tensors (float64);
    returns a float64 CUDA tensor when any input is a CUDA tensor, else a NumPy array."""
    arrays = (xs, ys, zs, xg, yg, zg, gas_mass, gas_Z, gas_h)
    on_device = any(dv.is_device_tensor(a) for a in arrays)
    d = [dv.to_dev(a) for a in arrays]
    so = np.ascontiguousarray(np.asarray(star_offsets.cpu() if isinstance(star_offsets, torch.Tensor) else star_offsets), dtype=np.int64)
    go = np.ascontiguousarray(np.asarray(gas_offsets.cpu() if isinstance(gas_offsets, torch.Tensor) else gas_offsets), dtype=np.int64)
    if so.ndim != 1 or so.shape != go.shape or so.size < 2:
        raise ValueError('star_offsets and gas_offsets must be CSR offset arrays of the same length n_obj + 1')
    n_obj = so.size - 1
    if so[0] != 0 or go[0] != 0 or so[-1] != d[0].numel() or go[-1] != d[3].numel():
        raise ValueError('offsets must start at 0 and end at the number of star / gas particles')
    if not (d[0].numel() == d[1].numel() == d[2].numel()) or any(a.numel() != d[3].numel() for a in d[4:]):
        raise ValueError('star arrays and gas arrays must have one entry per particle')
    _, table = kernel_table()
    out = torch.zeros(d[0].numel(), dtype=torch.float64, device=dv.device())
    wsb = _cabi.lib.so_los_workspace_bytes(n_obj, d[0].numel())
    ws = dv.workspace(wsb)
    i64p = ctypes.POINTER(ctypes.c_int64)
    _cabi.check(_cabi.lib.so_los_column(_cabi.ptr(d[0]), _cabi.ptr(d[1]), _cabi.ptr(d[2]), so.ctypes.data_as(i64p),
                                        _cabi.ptr(d[3]), _cabi.ptr(d[4]), _cabi.ptr(d[5]), _cabi.ptr(d[6]), _cabi.ptr(d[7]),
                                        _cabi.ptr(d[8]), go.ctypes.data_as(i64p), n_obj, _cabi.ptr(table), NT, float(scale),
                                        _cabi.ptr(out), _cabi.ptr(ws), wsb, _cabi.stream()), 'so_los_column')
    return out if on_device else dv.to_host(out)


GRID_MIN_PAIRS = 5.0e9        # star x gas pairs of one object from which the cell grid pays (building it costs ~2 ms =
                              # what all pairs cost at ~4e9; measured on B200)
GRID_CELL = 4.0               # cell edge in units of the median smoothing length (1e5 x 1e5, h = 0.05..1: kernel 2.3 / 2.3 /
                              # 1.85 / 1.44 ms at 0.5 / 1 / 2 / 4: the widest kernels sit in every list whatever the cell)
GRID_MAX_ENTRIES = 4.0e7      # (cell, gas) entries the grid may create (the gas arrays are gathered once per entry)


def _cell_grid(xs, ys, xg, yg, hg, max_cells=512):
    """Uniform grid over the stars' bounding box (cell edge ~ GRID_CELL median h): the stars sorted by cell and, for every
    cell, the list of the gas particles whose kernel footprint [x - h, x + h] x [y - h, y + h] meets it, in gas order
    (stable sorts: deterministic).  A star can only lie under kernels of its cell's list, so every cell is one independent
    'object' of so_los_column (cells without stars cost nothing there).  Returns (star order, star offsets, gas index per
    entry, gas offsets), the offsets over ALL n x n cells as host arrays, or None when the entry budget is exceeded (the
    caller then takes all pairs).  Three host synchronisations: bounds, number of entries, offsets."""
    dev = xs.device
    b = torch.stack([xs.min(), ys.min(), xs.max(), ys.max(), hg.median()]).cpu().tolist()
    x0, y0 = b[0], b[1]
    ext = max(b[2] - x0, b[3] - y0, 0.0)
    cell = max(GRID_CELL * b[4], ext / max_cells, 1e-300)      # stars inside one kernel width: a single cell
    for _ in range(6):
        n = int(np.floor(ext / cell)) + 1
        gx0, gx1 = torch.floor((xg - hg - x0) / cell).long(), torch.floor((xg + hg - x0) / cell).long()
        gy0, gy1 = torch.floor((yg - hg - y0) / cell).long(), torch.floor((yg + hg - y0) / cell).long()
        ok = (gx1 >= 0) & (gx0 <= n - 1) & (gy1 >= 0) & (gy0 <= n - 1)
        gx0, gx1, gy0, gy1 = (t.clamp(0, n - 1) for t in (gx0, gx1, gy0, gy1))
        nx = torch.where(ok, gx1 - gx0 + 1, torch.zeros_like(gx0))
        cnt = nx * torch.where(ok, gy1 - gy0 + 1, torch.zeros_like(gy0))
        total = int(cnt.sum())
        if total <= GRID_MAX_ENTRIES:
            break
        cell *= 2.0
    else:
        return None
    ncell = n * n
    sc = ((xs - x0) / cell).long().clamp(0, n - 1) + n * ((ys - y0) / cell).long().clamp(0, n - 1)
    order = torch.argsort(sc, stable=True)
    star_cnt = torch.zeros(ncell, dtype=torch.int64, device=dev).index_add_(0, sc, torch.ones_like(sc))
    gi = torch.repeat_interleave(torch.arange(xg.numel(), device=dev), cnt, output_size=total)
    k = torch.arange(total, device=dev) - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt, output_size=total)
    nxe = nx[gi]
    ce = (gx0[gi] + k % nxe) + n * (gy0[gi] + k // nxe)
    gi = gi[torch.argsort(ce, stable=True)]
    gas_cnt = torch.zeros(ncell, dtype=torch.int64, device=dev).index_add_(0, ce, torch.ones_like(ce))
    zero = torch.zeros(1, dtype=torch.int64, device=dev)
    off = torch.cat([zero, torch.cumsum(star_cnt, 0), zero, torch.cumsum(gas_cnt, 0)]).cpu().numpy()
    return order, np.ascontiguousarray(off[:ncell + 1]), gi, np.ascontiguousarray(off[ncell + 1:])


def metal_surface_density(xs, ys, zs, xg, yg, zg, gas_mass, gas_Z, gas_h, scale=1.0, grid='auto'):
    """Sigma_i for the star particles (xs, ys, zs) of ONE object from its gas particles (positions, masses, metallicities,
    smoothing lengths = kernel support radii), line of sight along +z.  scale = 1: units of mass / length^2 of the inputs
    (M_sol / kpc^2 for kpc and M_sol); 1e-6 gives M_sol / pc^2.
    grid: True / False / 'auto' (from GRID_MIN_PAIRS star x gas pairs on): bin the gas into the cells of a uniform grid
    over the stars and let every star test only the gas whose footprint meets its cell, instead of all of it.  Same
    specification, same pair arithmetic; only the float64 summation order over the tiles differs (~1e-16)."""
    ns = int(np.asarray(xs.shape if hasattr(xs, 'shape') else (len(xs),))[0])
    ng = int(np.asarray(xg.shape if hasattr(xg, 'shape') else (len(xg),))[0])
    use = (float(ns) * ng >= GRID_MIN_PAIRS) if grid == 'auto' else bool(grid)
    if use and ns > 0 and ng > 0:
        arrays = (xs, ys, zs, xg, yg, zg, gas_mass, gas_Z, gas_h)
        on_device = any(dv.is_device_tensor(a) for a in arrays)
        d = [dv.to_dev(a) for a in arrays]
        plan = _cell_grid(d[0], d[1], d[3], d[4], d[8])
        if plan is not None:
            order, so, gi, go = plan
            res = metal_surface_density_batch(d[0][order], d[1][order], d[2][order], so, *(a[gi] for a in d[3:]), go, scale=scale)
            out = torch.empty_like(res)
            out[order] = res
            return out if on_device else dv.to_host(out)
    return metal_surface_density_batch(xs, ys, zs, np.array([0, ns]), xg, yg, zg, gas_mass, gas_Z, gas_h,
                                       np.array([0, ng]), scale=scale)
