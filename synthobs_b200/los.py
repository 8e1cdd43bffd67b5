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


def metal_surface_density(xs, ys, zs, xg, yg, zg, gas_mass, gas_Z, gas_h, scale=1.0):
    """Sigma_i for the star particles (xs, ys, zs) of ONE object from its gas particles (positions, masses, metallicities,
    smoothing lengths = kernel support radii), line of sight along +z.  scale = 1: units of mass / length^2 of the inputs
    (M_sol / kpc^2 for kpc and M_sol); 1e-6 gives M_sol / pc^2."""
    ns = int(np.asarray(xs.shape if hasattr(xs, 'shape') else (len(xs),))[0])
    ng = int(np.asarray(xg.shape if hasattr(xg, 'shape') else (len(xg),))[0])
    return metal_surface_density_batch(xs, ys, zs, np.array([0, ns]), xg, yg, zg, gas_mass, gas_Z, gas_h,
                                       np.array([0, ng]), scale=scale)
