"""synthobs.sed.models — same public API as the reference module
(reference: synthobs/sed/models.py), computed by the sm_100a kernels behind
include/synthobs_b200.h.  No CPU path: every particle loop of the reference is a
CUDA kernel here; the host only prepares small float64 tables (filter weights,
dust-curve values, NGP thresholds) exactly as the reference's NumPy would.

Same names, positional order, defaults and return shapes as the reference:
define_model(.create_Lnu_grid, .create_Fnu_grid), generate_Lnu, generate_Lnu_array,
generate_Fnu, generate_Fnu_array, generate_log10Q, generate_SED.
Additions (no upstream equivalent): generate_Lnu_batch /
generate_Fnu_batch take CSR `offsets` and return [n_galaxies, n_filters].

Inputs may be NumPy arrays (as in the reference), CPU tensors (pinned for fast
H2D) or CUDA tensors; host inputs give NumPy float64 outputs, CUDA inputs give
CUDA tensors.
"""

import copy
import os
import pickle
from ctypes import c_double, c_int64

import numpy as np
import torch

from .. import _cabi
from .. import device as dv
from .. import providers

core, IGM, dust_curves = providers.resolve()

try:  # the reference takes its default grid location from flare (synthobs/core.py:6-8)
    from flare import FLARE_dir  # type: ignore
except Exception:  # offline: $FLARE or empty
    FLARE_dir = os.environ.get('FLARE', '')

COMPONENTS = ('stellar_incident', 'stellar_transmitted', 'nebular')


class empty():
    pass


# --------------------------------------------------------------------------
# NGP thresholds: bit-exact np.argmin without a device log10
# --------------------------------------------------------------------------

def _bisect_bits(pred, m):
    """Per element i of a length-m problem: the smallest positive float64 v with
    pred(v)[i] True, for predicates monotone in v (False ... False True ... True).
    Bisection on the bit patterns (positive doubles order like their uint64 bits).
    Returns +inf where no finite value satisfies the predicate."""
    lo = np.full(m, 1, dtype=np.uint64)                                   # smallest denormal
    hi = np.full(m, np.array(np.inf).view(np.uint64), dtype=np.uint64)    # bits(+inf) = "none"
    while np.any(lo < hi):
        mid = lo + (hi - lo) // np.uint64(2)
        with np.errstate(all='ignore'):
            ok = pred(mid.view(np.float64))
        hi = np.where(ok, mid, hi)
        lo = np.where(ok, lo, mid + np.uint64(1))
    return hi.view(np.float64)


def ngp_thresholds(nodes, offset):
    """thr[i] = smallest v > 0 for which the reference's
    argmin|nodes - (log10(v) + offset)| (models.py:108-114) is >= i+1.
    The index is then  #{i : v >= thr[i]}  — float64 compares only.
    The node axis must be strictly ascending (every SSP grid is); np.argmin itself would accept any
    order, the threshold form does not, so anything else is refused instead of silently mis-indexed."""
    nodes = np.asarray(nodes, dtype=np.float64)
    n = nodes.size
    if n < 2:
        return np.zeros(0)
    if not (np.all(np.isfinite(nodes)) and np.all(np.diff(nodes) > 0)):
        raise ValueError('the SSP grid axes (log10age, log10Z) must be finite and strictly ascending for the '
                         'exact nearest-grid-point thresholds; got %r' % (nodes,))
    want = np.arange(1, n)

    def index_of(v):
        with np.errstate(all='ignore'):
            val = np.log10(v) + offset
            return np.abs(nodes[None, :] - val[:, None]).argmin(axis=1)

    thr = _bisect_bits(lambda v: index_of(v) >= want, n - 1)
    # verify the monotonicity assumption a few ulps around every threshold
    fin = np.isfinite(thr)
    up, dn = thr.copy(), np.nextafter(thr, 0.0)
    for _ in range(4):
        if np.any(index_of(up)[fin] < want[fin]) or np.any(index_of(dn)[fin & (dn > 0)] >= want[fin & (dn > 0)]):
            raise RuntimeError('NGP index is not monotone around a threshold; cannot build exact thresholds')
        up, dn = np.nextafter(up, np.inf), np.nextafter(dn, 0.0)
    # and end to end on a sample spanning the axis: the reference's argmin == the number of thresholds passed
    lo, hi = nodes[0] - offset - 1.0, nodes[-1] - offset + 1.0
    sample = 10.0 ** np.linspace(lo, hi, 4096)
    if not np.array_equal(index_of(sample), (sample[:, None] >= thr[None, :]).sum(axis=1)):
        raise RuntimeError('NGP thresholds do not reproduce argmin on the sample; cannot build exact thresholds')
    return thr


def young_threshold(log10t_BC):
    """Smallest Age [Myr] that is NOT young, i.e. not (log10(Age)+6 < log10t_BC)
    (models.py:124)."""
    return float(_bisect_bits(lambda v: ~((np.log10(v) + 6.) < log10t_BC), 1)[0])


def trapz_weights(x):
    """w with np.trapz(y, x=x) == sum(w*y) (exact identity)."""
    x = np.asarray(x, dtype=np.float64)
    w = np.zeros_like(x)
    d = np.diff(x)
    w[:-1] += 0.5 * d
    w[1:] += 0.5 * d
    return w


# --------------------------------------------------------------------------
# model
# --------------------------------------------------------------------------

class _NGPGrid():
    """(log10age, log10Z) node axes of a grid dict + the device-side threshold caches
    shared by define_model and EmissionLines (both do the NGP look-up of models.py:108-114)."""

    @property
    def na(self):
        return int(self.grid['log10age'].size)

    @property
    def nz(self):
        return int(self.grid['log10Z'].size)

    def _thresholds(self):
        c = self._dev.get('thr')
        if c is None:
            la = np.asarray(self.grid['log10age'], dtype=np.float64)
            age_thr = ngp_thresholds(la, 6.0)
            z_thr = ngp_thresholds(self.grid['log10Z'], 0.0)
            c = {'age': _cabi.host_doubles(age_thr), 'z': _cabi.host_doubles(z_thr),
                 'age_host': age_thr, 'z_host': z_thr, 'young': {}}
            self._dev['thr'] = c
        return c

    def _young_thr(self, log10t_BC):
        c = self._thresholds()['young']
        key = float(log10t_BC)
        if key not in c:
            c[key] = young_threshold(key)
        return c[key]

    def invalidate(self):
        """Drop every device-side cache (thresholds, filter tables, SSP-grid operands, dust-curve values).
        The reference reads model.grid / model.Lnu / model.Fnu / F[f].T at call time; here they are uploaded
        once and keyed by object identity, so call this after editing any of them IN PLACE (replacing the
        objects, e.g. by create_Lnu_grid, is noticed without it)."""
        self._dev.clear()


class define_model(_NGPGrid):
    """SSP grid container (reference: synthobs/sed/models.py:17-67)."""

    def __init__(self, grid, path_to_SPS_grid=FLARE_dir + '/data/SPS/nebular/3.0/'):

        if isinstance(grid, dict):   # convenience for synthetic grids: pass the dict itself
            self.grid = grid
        else:
            self.grid = pickle.load(open(path_to_SPS_grid + grid + '/nebular.p', 'rb'), encoding='latin1')

        self.lam = self.grid['lam']

        self.dust_ISM = False
        self.dust_BC = False

        print(list(self.grid.keys()))

        if 'stellar_incident' not in self.grid:
            # canonical derivation (models.py:30-35)
            self.grid['stellar_incident'] = copy.copy(self.grid['stellar'])
            self.grid['stellar_transmitted'] = copy.copy(self.grid['stellar'])
            self.grid['stellar_transmitted'][:, :, self.lam < 912] = 0.0

        self._dev = {}

    # ---- device-side caches -------------------------------------------------
    def _grid_rows(self):
        """[3*ncell, n_lam] float32 on the device (incident, transmitted, nebular
        rows), wavelength contiguous: the K-major left operand of the grid x
        filter-weights contraction."""
        c = self._dev.get('rows')
        if c is None:
            ncell = self.na * self.nz
            nl = self.lam.size
            nlp = (nl + 3) // 4 * 4
            rows = np.zeros((3 * ncell, nlp), dtype=np.float32)
            for i, comp in enumerate(COMPONENTS):
                rows[i * ncell:(i + 1) * ncell, :nl] = np.asarray(self.grid[comp]).reshape(ncell, nl)
            c = dv.to_dev(rows, torch.float32)
            self._dev['rows'] = c
        return c

    def _contract_with_filters(self, Wt):
        """[3*ncell, n_lam] x Wt[F, n_lam]^T -> float64 [3, na, nz, F] (K4)."""
        _cabi.require_device()
        rows = self._grid_rows()
        nlp = rows.shape[1]
        F = Wt.shape[0]
        W = np.zeros((F, nlp), dtype=np.float32)
        W[:, :Wt.shape[1]] = Wt
        Wd = dv.to_dev(W, torch.float32)
        out = gemm(rows, Wd)
        return dv.to_host(out).reshape(3, self.na, self.nz, F)

    # ---- reference API -------------------------------------------------------
    def create_Lnu_grid(self, F):
        """reference: models.py:39-49.  trapz(grid*T, lam)/trapz(T, lam) for every
        filter and component as ONE dense contraction grid[3*ncell, n_lam] x
        W[n_lam, F], W = trapz weights * T / norm folded in float64 on the host."""
        filters = F['filters']
        Wt = np.stack([trapz_weights(F[f].lam) * F[f].T / np.trapezoid(F[f].T, x=F[f].lam) for f in filters])
        scale = np.abs(Wt).max(axis=1, keepdims=True)
        scale[scale == 0] = 1.0
        tabs = self._contract_with_filters(Wt / scale) * scale[:, 0]
        self.Lnu = {}
        for fi, f in enumerate(filters):
            self.Lnu[f] = empty()
            for ci, comp in enumerate(COMPONENTS):
                setattr(self.Lnu[f], comp, np.ascontiguousarray(tabs[ci, :, :, fi]))

    def create_Fnu_grid(self, F, z, cosmo):
        """reference: models.py:53-67 (nJy per Msol).  IGM transmission and the
        flux scaling 1e23*1e9*(1+z)/(4 pi dL^2) are folded into the weights in
        float64 (the scale itself does not fit FP32 comfortably)."""
        self.Fnu = {}
        self.Fnu['z'] = z
        luminosity_distance = cosmo.luminosity_distance(z).to('cm').value
        filters = F['filters']
        flux = 1E23 * 1E9 * (1. + z) / (4. * np.pi * luminosity_distance ** 2)
        Wt = np.stack([trapz_weights(F[f].lam) * F[f].T * IGM.madau(F[f].lam, z) / np.trapezoid(F[f].T, x=F[f].lam)
                       for f in filters])
        scale = np.abs(Wt).max(axis=1, keepdims=True)
        scale[scale == 0] = 1.0
        tabs = self._contract_with_filters(Wt / scale) * (scale[:, 0] * flux)
        for fi, f in enumerate(filters):
            self.Fnu[f] = empty()
            for ci, comp in enumerate(COMPONENTS):
                setattr(self.Fnu[f], comp, np.ascontiguousarray(tabs[ci, :, :, fi]))


# --------------------------------------------------------------------------
# dense contraction helper (K4/K5)
# --------------------------------------------------------------------------

GEMM_ALGO = 'auto'   # 'auto' | '3xtf32' | 'f32'


def gemm(A, B, algo=None, B_split=None):
    """C[M,N] = A[M,K] @ B[N,K]^T on the device (float32 in/out).  '3xtf32' is the
    tcgen05 tensor-core kernel with error compensation (default); 'f32' the
    FMA-pipe kernel kept for on-device validation."""
    _cabi.require_device()
    M, K = A.shape
    N = B.shape[0]
    algo = algo or GEMM_ALGO
    if algo == 'auto':
        algo = '3xtf32'
    C = torch.empty((M, N), dtype=torch.float32, device=A.device)
    if algo == '3xtf32':
        Kp = (K + 3) // 4 * 4       # TMA needs 16-byte row strides

        def prep(X):
            if X.stride(1) == 1 and X.stride(0) % 4 == 0 and X.stride(0) >= K and X.data_ptr() % 16 == 0:
                return X
            Y = torch.zeros((X.shape[0], Kp), dtype=torch.float32, device=X.device)
            Y[:, :K] = X
            return Y
        A = prep(A)
        Ah, Al = split_tf32(A)
        if B_split is not None:      # static operand (the SSP grid): split once per model
            Bh, Bl = B_split
            B = Bh
        else:
            B = prep(B)
            Bh, Bl = split_tf32(B)
        rc = _cabi.lib.so_gemm_3xtf32(_cabi.ptr(Ah), _cabi.ptr(Al), A.stride(0), _cabi.ptr(Bh), _cabi.ptr(Bl),
                                      B.stride(0), _cabi.ptr(C), C.stride(0), M, N, K, _cabi.stream())
        _cabi.check(rc, 'so_gemm_3xtf32')
    else:
        A, B = A.contiguous(), B.contiguous()
        rc = _cabi.lib.so_gemm_f32(_cabi.ptr(A), A.stride(0), _cabi.ptr(B), B.stride(0), _cabi.ptr(C), C.stride(0),
                                   M, N, K, _cabi.stream())
        _cabi.check(rc, 'so_gemm_f32')
    return C


def split_tf32(x):
    """x (dense rows, possibly with a padded row stride) -> TF32-exact hi and lo parts"""
    n = x.stride(0) * (x.shape[0] - 1) + x.shape[1] if x.dim() == 2 else x.numel()
    hi = torch.empty_strided(x.shape, x.stride(), dtype=x.dtype, device=x.device)
    lo = torch.empty_strided(x.shape, x.stride(), dtype=x.dtype, device=x.device)
    _cabi.check(_cabi.lib.so_split_tf32(_cabi.ptr(x), _cabi.ptr(hi), _cabi.ptr(lo), n, _cabi.stream()),
                'so_split_tf32')
    return hi, lo


# --------------------------------------------------------------------------
# broadband (K1+K3)
# --------------------------------------------------------------------------

def _dust_k(model, which, wavelength):
    spec = model.dust_ISM if which == 'ISM' else model.dust_BC
    if not spec:
        return None
    dust_model, dust_model_params = spec
    return getattr(dust_curves, dust_model)(params=dust_model_params).tau(wavelength)   # models.py:97


def _pow2_scale(*arrays):
    """Per-row power of two >= max|value| when that maximum would leave the FP32 range once
    multiplied by a particle mass (line luminosities per Msol reach 1e34 erg/s); 1 otherwise.
    Dividing by a power of two is exact in binary floating point."""
    m = np.max(np.stack([np.abs(a).max(axis=1) for a in arrays]), axis=0)
    need = (m > 1e24) | ((m < 1e-24) & (m > 0))
    with np.errstate(divide='ignore'):
        e = np.where(need, np.ceil(np.log2(np.where(m > 0, m, 1.0))), 0.0)
    return np.ldexp(1.0, e.astype(np.int64))


def _scaled_tables(inc, tn):
    """float64 [F, ncell] tables -> (device float32 inc/scale, tn/scale, device float64 scale [F] or
    None when no row needs scaling)"""
    scale = _pow2_scale(inc, tn)
    return (dv.to_dev((inc / scale[:, None]).astype(np.float32), torch.float32),
            dv.to_dev((tn / scale[:, None]).astype(np.float32), torch.float32),
            dv.to_dev(scale, torch.float64) if np.any(scale != 1.0) else None)


def _tables(model, kind, filters):
    """Device tables [F, ncell] float32 built from whatever model.Lnu / model.Fnu
    currently hold (the reference reads them at call time, models.py:131-133)."""
    src = model.Lnu if kind == 'Lnu' else model.Fnu
    arrays = [getattr(src[f], c) for f in filters for c in COMPONENTS]
    key = (kind, tuple(filters), tuple(id(a) for a in arrays))
    cache = model._dev.setdefault('tabs', {})
    if key not in cache:
        # the cache keeps the keyed arrays alive, so an id() cannot be recycled while its entry exists; arrays
        # REPLACED on model.Lnu / model.Fnu are picked up at the next call, arrays edited IN PLACE need
        # model.invalidate()
        model._dev['tabs_keep'] = arrays
        inc = np.stack([np.asarray(src[f].stellar_incident, dtype=np.float64).reshape(-1) for f in filters])
        tn = np.stack([(np.asarray(src[f].stellar_transmitted, dtype=np.float64)
                        + np.asarray(src[f].nebular, dtype=np.float64)).reshape(-1) for f in filters])
        cache.clear()
        cache[key] = _scaled_tables(inc, tn)
    return cache[key]


def _opt(x, like):
    if isinstance(x, (np.ndarray, torch.Tensor)):
        return dv.to_dev(x)
    return None     # reference: missing tauV arrays are zeros (models.py:92-93)


def _broadband(model, kind, F, filters, Masses, Ages, Metallicities, tauVs_ISM, tauVs_BC, fesc, log10t_BC,
               offsets=None, want_pf=False, want_sums=True, want_cell=False):
    _cabi.require_device()
    tabs = _tables(model, kind, filters)
    zshift = (1. + model.Fnu['z']) if kind == 'Fnu' else 1.0                    # models.py:163,167
    ckey = (kind, id(F), tuple(filters), repr(model.dust_ISM), repr(model.dust_BC), zshift)
    kcache = model._dev.setdefault('dustk', {})
    if ckey not in kcache:     # pivot wavelengths cost two trapz per filter: do them once per (F, dust) setup
        piv = [F[f].pivwv() / zshift for f in filters]
        kcache.clear()
        model._dev['dustk_keep'] = F     # keeps id(F) from being recycled while the entry exists
        kcache[ckey] = (None if not model.dust_ISM else [float(_dust_k(model, 'ISM', p)) for p in piv],
                        None if not model.dust_BC else [float(_dust_k(model, 'BC', p)) for p in piv])
    kI, kB = kcache[ckey]
    return _broadband_launch(model, tabs, kI, kB, Masses, Ages, Metallicities, tauVs_ISM, tauVs_BC, fesc, log10t_BC,
                             offsets=offsets, want_pf=want_pf, want_sums=want_sums, want_cell=want_cell)


def _broadband_launch(ngp, tabs, kI, kB, Masses, Ages, Metallicities, tauVs_ISM, tauVs_BC, fesc, log10t_BC,
                      offsets=None, want_pf=False, want_sums=True, want_cell=False):
    """One so_sed_broadband pass.  `ngp` supplies the grid axes (_NGPGrid), `tabs` =
    (inc/scale, tn/scale, scale) device tables [F, ncell], kI / kB = per-table dust-curve
    values (lists) or None.  Returns (per-particle [F, n] float32 | None, sums [n_seg, F]
    float64 | None, cell ids | None), already multiplied back by `scale`."""
    tab_inc, tab_tn, scale = tabs
    M = dv.to_dev(Masses)
    A = dv.to_dev(Ages)
    Z = dv.to_dev(Metallicities)
    tI = _opt(tauVs_ISM, M)
    tB = _opt(tauVs_BC, M)
    n = M.numel()
    nf = tab_inc.shape[0]
    thr = ngp._thresholds()
    if offsets is not None:
        off = dv.to_dev(offsets, torch.int64)
        nseg = off.numel() - 1
    else:
        off, nseg = None, 1
    dev = M.device
    out_pf = torch.empty((nf, n), dtype=torch.float32, device=dev) if want_pf else None
    out_sums = torch.empty((nseg, nf), dtype=torch.float64, device=dev) if want_sums else None
    out_cell = torch.empty(n, dtype=torch.int32, device=dev) if want_cell else None
    wsb = _cabi.lib.so_sed_broadband_workspace_bytes(n, nseg, nf)
    ws = dv.workspace(wsb)
    rc = _cabi.lib.so_sed_broadband(
        _cabi.ptr(M), _cabi.ptr(A), _cabi.ptr(Z), _cabi.ptr(tI), _cabi.ptr(tB), n,
        _cabi.ptr(off), nseg,
        thr['age'], ngp.na, thr['z'], ngp.nz, ngp._young_thr(log10t_BC),
        _cabi.ptr(tab_inc), _cabi.ptr(tab_tn), nf,
        _cabi.host_doubles(kI), _cabi.host_doubles(kB), float(fesc),
        _cabi.ptr(out_pf), _cabi.ptr(out_sums), _cabi.ptr(out_cell),
        _cabi.ptr(ws), wsb, _cabi.stream())
    _cabi.check(rc, 'so_sed_broadband')
    if scale is not None:
        if out_sums is not None:
            out_sums *= scale[None, :]
        if out_pf is not None:
            out_pf *= scale.to(torch.float32)[:, None]
    return out_pf, out_sums, out_cell


def _host_like(*xs):
    return not any(dv.is_device_tensor(x) for x in xs)


def generate_Lnu(model, F, Masses, Ages, Metallicities, tauVs_ISM=False, tauVs_BC=False, fesc=0.0, log10t_BC=7.):
    """reference: models.py:74-85 — {filter: total Lnu}; all filters in one kernel pass."""
    _, sums, _ = _broadband(model, 'Lnu', F, F['filters'], Masses, Ages, Metallicities, tauVs_ISM, tauVs_BC,
                            fesc, log10t_BC)
    s = sums[0].cpu().numpy()
    return {f: s[i] for i, f in enumerate(F['filters'])}


def generate_Lnu_array(model, F, f, Masses, Ages, Metallicities, tauVs_ISM=False, tauVs_BC=False, fesc=0.0,
                       log10t_BC=7.):
    """reference: models.py:88-135 — per-particle Lnu in filter f."""
    pf, _, _ = _broadband(model, 'Lnu', F, [f], Masses, Ages, Metallicities, tauVs_ISM, tauVs_BC, fesc, log10t_BC,
                          want_pf=True, want_sums=False)
    return dv.to_host(pf[0]) if _host_like(Masses) else pf[0]


def generate_Fnu(model, F, Masses, Ages, Metallicities, tauVs_ISM=False, tauVs_BC=False, fesc=0.0, log10t_BC=7.):
    """reference: models.py:139-150."""
    _, sums, _ = _broadband(model, 'Fnu', F, F['filters'], Masses, Ages, Metallicities, tauVs_ISM, tauVs_BC,
                            fesc, log10t_BC)
    s = sums[0].cpu().numpy()
    return {f: s[i] for i, f in enumerate(F['filters'])}


def generate_Fnu_array(model, F, f, Masses, Ages, Metallicities, tauVs_ISM=False, tauVs_BC=False, fesc=0.0,
                       log10t_BC=7.):
    """reference: models.py:153-200."""
    pf, _, _ = _broadband(model, 'Fnu', F, [f], Masses, Ages, Metallicities, tauVs_ISM, tauVs_BC, fesc, log10t_BC,
                          want_pf=True, want_sums=False)
    return dv.to_host(pf[0]) if _host_like(Masses) else pf[0]


def generate_Lnu_arrays(model, F, Masses, Ages, Metallicities, tauVs_ISM=False, tauVs_BC=False, fesc=0.0,
                        log10t_BC=7., kind='Lnu'):
    """All filters' per-particle arrays in one pass: {filter: [N]} (what the imaging
    examples build with a dict comprehension over generate_Fnu_array,
    examples/morph_image_observed.py:93)."""
    pf, _, _ = _broadband(model, kind, F, F['filters'], Masses, Ages, Metallicities, tauVs_ISM, tauVs_BC, fesc,
                          log10t_BC, want_pf=True, want_sums=False)
    host = _host_like(Masses)
    return {f: (dv.to_host(pf[i]) if host else pf[i]) for i, f in enumerate(F['filters'])}


def generate_Fnu_arrays(model, F, Masses, Ages, Metallicities, **kw):
    return generate_Lnu_arrays(model, F, Masses, Ages, Metallicities, kind='Fnu', **kw)


def generate_Lnu_batch(model, F, Masses, Ages, Metallicities, offsets, tauVs_ISM=False, tauVs_BC=False, fesc=0.0,
                       log10t_BC=7., kind='Lnu'):
    """Many galaxies in one call: particles concatenated, `offsets` = CSR [G+1].
    Equals a Python loop of generate_Lnu over the galaxies.  Returns [G, F]
    (NumPy for host inputs, CUDA tensor otherwise)."""
    _, sums, _ = _broadband(model, kind, F, F['filters'], Masses, Ages, Metallicities, tauVs_ISM, tauVs_BC, fesc,
                            log10t_BC, offsets=offsets)
    return sums.cpu().numpy() if _host_like(Masses) else sums


def generate_Fnu_batch(model, F, Masses, Ages, Metallicities, offsets, **kw):
    return generate_Lnu_batch(model, F, Masses, Ages, Metallicities, offsets, kind='Fnu', **kw)


def ngp_cells(model, Ages, Metallicities, log10t_BC=7.):
    """(ia, iZ, young) as the kernels see them — for index-parity tests."""
    _cabi.require_device()
    A = dv.to_dev(Ages)
    Z = dv.to_dev(Metallicities)
    n = A.numel()
    thr = model._thresholds()
    ncell = model.na * model.nz
    dummy = torch.zeros((1, ncell), dtype=torch.float32, device=A.device)
    ones = torch.ones(n, dtype=torch.float64, device=A.device)
    cell = torch.empty(n, dtype=torch.int32, device=A.device)
    wsb = _cabi.lib.so_sed_broadband_workspace_bytes(n, 1, 1)
    ws = dv.workspace(wsb)
    rc = _cabi.lib.so_sed_broadband(
        _cabi.ptr(ones), _cabi.ptr(A), _cabi.ptr(Z), None, None, n, None, 1,
        thr['age'], model.na, thr['z'], model.nz, model._young_thr(log10t_BC),
        _cabi.ptr(dummy), _cabi.ptr(dummy), 1, None, None, 0.0,
        None, None, _cabi.ptr(cell), _cabi.ptr(ws), wsb, _cabi.stream())
    _cabi.check(rc, 'so_sed_broadband')
    c = cell.cpu().numpy()
    code = c & 65535
    return code // model.nz, code % model.nz, (c >> 16).astype(bool)


# --------------------------------------------------------------------------
# histogram-based quantities (K2, K5)
# --------------------------------------------------------------------------

def _cells_and_hist(model, Masses, Ages, Metallicities, log10t_BC, offsets=None, split=False):
    _cabi.require_device()
    M = dv.to_dev(Masses)
    A = dv.to_dev(Ages)
    Z = dv.to_dev(Metallicities)
    n = M.numel()
    thr = model._thresholds()
    ncell = model.na * model.nz
    dummy = torch.zeros((1, ncell), dtype=torch.float32, device=M.device)
    cell = torch.empty(n, dtype=torch.int32, device=M.device)
    if offsets is not None:
        off = dv.to_dev(offsets, torch.int64)
        nseg = off.numel() - 1
    else:
        off, nseg = None, 1
    wsb = _cabi.lib.so_sed_broadband_workspace_bytes(n, nseg, 1)
    ws = dv.workspace(wsb)
    rc = _cabi.lib.so_sed_broadband(
        _cabi.ptr(M), _cabi.ptr(A), _cabi.ptr(Z), None, None, n, _cabi.ptr(off), nseg,
        thr['age'], model.na, thr['z'], model.nz, model._young_thr(log10t_BC),
        _cabi.ptr(dummy), _cabi.ptr(dummy), 1, None, None, 0.0,
        None, None, _cabi.ptr(cell), _cabi.ptr(ws), wsb, _cabi.stream())
    _cabi.check(rc, 'so_sed_broadband')
    ld = (2 * ncell + 3) // 4 * 4
    H = torch.zeros((nseg, ld), dtype=torch.float32, device=M.device)
    Hlo = torch.zeros_like(H) if split else None
    wsb = _cabi.lib.so_sed_hist_workspace_bytes(nseg, ncell)
    ws = dv.workspace(wsb)
    rc = _cabi.lib.so_sed_hist(_cabi.ptr(M), _cabi.ptr(cell), n, _cabi.ptr(off), nseg, ncell,
                               _cabi.ptr(H), _cabi.ptr(Hlo), ld, _cabi.ptr(ws), wsb, _cabi.stream())
    _cabi.check(rc, 'so_sed_hist')
    return M, cell, H, Hlo, ncell


def generate_log10Q(model, Masses, Ages, Metallicities):
    """reference: models.py:206-221 — log10 sum_i M_i 10^log10Q[ia,iZ] = log10(H . 10^log10Q)."""
    M, cell, H, _, ncell = _cells_and_hist(model, Masses, Ages, Metallicities, 7.0)
    h = H[0, :2 * ncell].to(torch.float64).cpu().numpy()
    h = h[:ncell] + h[ncell:]
    return np.log10(np.sum(h * 10 ** np.asarray(model.grid['log10Q'], dtype=np.float64).reshape(-1)))


class generate_SED():
    """reference: models.py:229-314.  stellar / intrinsic are histogram x grid
    contractions (K2 + K5, tensor cores); BC / total / no_HII_no_BC need
    exp(-tauV_i k(lambda)) per particle and wavelength (K6)."""

    def __init__(self, model, Masses, Ages, Metallicities, tauVs_ISM=False, tauVs_BC=False, IGM=False, fesc=0.0,
                 log10t_BC=7.):

        self.model = model
        self.lam = self.model.grid['lam']

        self.stellar = core.sed(self.lam, description='intrinsic (stellar) SED')
        self.intrinsic = core.sed(self.lam, description='SED with nebular emission (HII) but no BC or ISM dust')
        self.BC = core.sed(self.lam, description='SED with nebular emission (HII), BC dust, but no ISM dust')
        self.no_HII_no_BC = core.sed(self.lam, description='SED with no nebular emission (HII) or dusty BC (same as setting fesc=1.0)')
        self.total = core.sed(self.lam, description='SED including both birth cloud and ISM dust')

        if isinstance(tauVs_ISM, np.ndarray) and not self.model.dust_ISM:
            print("WARNING! You have supplied ISM optical depths but have not specificied a ISM dust model. This doesn't sound right!")
        if isinstance(tauVs_BC, np.ndarray) and not self.model.dust_BC:
            print("WARNING! You have supplied BC optical depths but have not specificied a BC dust model. This doesn't sound right!")

        spectra = sed_spectra_device(model, Masses, Ages, Metallicities, tauVs_ISM, tauVs_BC, fesc, log10t_BC)
        for k in ('stellar', 'intrinsic', 'BC', 'total', 'no_HII_no_BC'):
            getattr(self, k).lnu = dv.to_host(spectra[k])

        with np.errstate(divide='ignore', invalid='ignore'):
            self.f_esc = self.total.lnu / self.intrinsic.lnu                      # models.py:312
            self.tau = -np.log(self.f_esc)                                        # models.py:313
            self.tau_relV = self.tau / np.interp(5500., self.lam, self.tau)       # models.py:314


def _spectral_operands(model):
    """Device operands for the hist x grid contraction: BT[n_lam, 2*ncell_pad]
    = [incident^T | (transmitted+nebular)^T], and the same rows for K6."""
    c = model._dev.get('BT')
    if c is None:
        ncell = model.na * model.nz
        nl = model.lam.size
        ld = (2 * ncell + 3) // 4 * 4
        inc = np.asarray(model.grid['stellar_incident'], dtype=np.float64).reshape(ncell, nl)
        tn = (np.asarray(model.grid['stellar_transmitted'], dtype=np.float64)
              + np.asarray(model.grid['nebular'], dtype=np.float64)).reshape(ncell, nl)
        BT = np.zeros((nl, ld), dtype=np.float32)
        BT[:, :ncell] = inc.T
        BT[:, ncell:2 * ncell] = tn.T
        c = {'BT': dv.to_dev(BT, torch.float32),
             'inc': dv.to_dev(inc.astype(np.float32), torch.float32),
             'tn': dv.to_dev(tn.astype(np.float32), torch.float32), 'ld': ld}
        c['BT_split'] = split_tf32(c['BT'])
        model._dev['BT'] = c
    return c


def sed_spectra_device(model, Masses, Ages, Metallicities, tauVs_ISM=False, tauVs_BC=False, fesc=0.0,
                       log10t_BC=7., offsets=None):
    """The five spectra of generate_SED as float32 CUDA tensors: [n_lam] for one
    galaxy, [G, n_lam] when CSR `offsets` are given (batch of galaxies)."""
    M, cell, H, Hlo_unused, ncell = _cells_and_hist(model, Masses, Ages, Metallicities, log10t_BC, offsets=offsets)
    ops = _spectral_operands(model)
    ld = ops['ld']
    G = H.shape[0]
    Hold = H[:, :ncell]
    Hyng = H[:, ncell:2 * ncell]
    A = torch.zeros((2 * G, ld), dtype=torch.float32, device=H.device)
    A[:G, :ncell] = Hold + Hyng                           # stellar = sum_i M_i inc            (models.py:277,306)
    A[G:, :ncell] = Hold + float(fesc) * Hyng             # intrinsic, incident part           (models.py:281,295)
    A[G:, ncell:2 * ncell] = (1.0 - float(fesc)) * Hyng   # intrinsic, transmitted+nebular part
    C = gemm(A, ops['BT'], B_split=ops.get('BT_split'))   # [2G, n_lam] on the tensor cores
    squeeze = (lambda t: t[0]) if offsets is None else (lambda t: t)
    out = {'stellar': squeeze(C[:G]), 'intrinsic': squeeze(C[G:])}
    have_ism = bool(model.dust_ISM)
    have_bc = bool(model.dust_BC)
    if not have_ism and not have_bc:
        out['BC'] = out['intrinsic'].clone()              # T = 1 everywhere                   (models.py:285-304)
        out['total'] = out['intrinsic'].clone()
        out['no_HII_no_BC'] = out['stellar'].clone()
        return out
    n = M.numel()
    nl = model.lam.size
    tI = _opt(tauVs_ISM, M) if have_ism else None
    tB = _opt(tauVs_BC, M) if have_bc else None
    kI = dv.to_dev(np.asarray(_dust_k(model, 'ISM', model.lam), dtype=np.float32), torch.float32) if have_ism else None
    kB = dv.to_dev(np.asarray(_dust_k(model, 'BC', model.lam), dtype=np.float32), torch.float32) if have_bc else None
    if offsets is None:
        off = None
        key, order = torch.sort(cell, stable=True)
    else:
        off = dv.to_dev(offsets, torch.int64)
        # galaxy of every particle by bisection of the offsets; (galaxy, young, cell) fits 31 bits for up to 2^14
        # galaxies, so the radix sort runs four 8-bit passes instead of eight
        gal = torch.searchsorted(off[1:], torch.arange(n, device=M.device, dtype=torch.int64), right=True)
        if G <= (1 << 14):
            _, order = torch.sort((gal << 17).to(torch.int32) + cell, stable=True)
        else:
            _, order = torch.sort((gal << 17) + cell.to(torch.int64), stable=True)
        key = cell[order]
    order = order.to(torch.int32)
    bc = torch.empty((G, nl), dtype=torch.float32, device=M.device)
    tot = torch.empty_like(bc)
    noh = torch.empty_like(bc)
    wsb = _cabi.lib.so_sed_dust_spectra_workspace_bytes(n, G, nl)
    ws = dv.workspace(wsb)
    rc = _cabi.lib.so_sed_dust_spectra(_cabi.ptr(M), _cabi.ptr(tI), _cabi.ptr(tB), _cabi.ptr(key), _cabi.ptr(order), n,
                                       _cabi.ptr(off), G,
                                       _cabi.ptr(ops['inc']), _cabi.ptr(ops['tn']), ncell, nl,
                                       _cabi.ptr(kI), _cabi.ptr(kB), float(fesc),
                                       _cabi.ptr(bc), _cabi.ptr(tot), _cabi.ptr(noh), _cabi.ptr(ws), wsb, _cabi.stream())
    _cabi.check(rc, 'so_sed_dust_spectra')
    out['BC'], out['total'], out['no_HII_no_BC'] = squeeze(bc), squeeze(tot), squeeze(noh)
    return out


def generate_SED_batch(model, Masses, Ages, Metallicities, offsets, tauVs_ISM=False, tauVs_BC=False, fesc=0.0,
                       log10t_BC=7.):
    """generate_SED for many galaxies in one call (no upstream equivalent):
    dict of [G, n_lam] arrays (NumPy float64 for host inputs, CUDA float32 tensors otherwise)
    equal to a Python loop of generate_SED over the galaxies."""
    out = sed_spectra_device(model, Masses, Ages, Metallicities, tauVs_ISM, tauVs_BC, fesc, log10t_BC, offsets=offsets)
    if _host_like(Masses):
        return {k: dv.to_host(v) for k, v in out.items()}
    return out


# --------------------------------------------------------------------------
# spectra -> broadband (the second contraction)
# --------------------------------------------------------------------------

def _filter_weight_operand(model, F, kind, z=None, cosmo=None, include_IGM=True):
    """W[F, n_lam] (float32, rows scaled to unit maximum, TF32-split) + float64 row scale, for
    core.sed.get_Lnu (kind 'Lnu': trapz(lnu T)/trapz(T) on the rest-frame grid) or
    get_fnu + get_Fnu (kind 'Fnu': nJy, IGM and flux scaling folded in; F on lam(1+z))."""
    key = (kind, id(F), tuple(F['filters']), z, include_IGM, id(cosmo))
    cache = model._dev.setdefault('specW', {})
    if key not in cache:
        filters = F['filters']
        Wt = np.stack([trapz_weights(F[f].lam) * F[f].T / np.trapezoid(F[f].T, x=F[f].lam) for f in filters])
        if kind == 'Fnu':
            dl = cosmo.luminosity_distance(z).to('cm').value
            Wt = Wt * (1E23 * 1E9 * (1. + z) / (4. * np.pi * dl ** 2))
            if include_IGM:
                Wt = Wt * np.stack([IGM.madau(F[f].lam, z) for f in filters])
        scale = np.abs(Wt).max(axis=1)
        scale[scale == 0] = 1.0
        nl = Wt.shape[1]
        nlp = (nl + 3) // 4 * 4
        Wp = np.zeros((len(filters), nlp), dtype=np.float32)
        Wp[:, :nl] = Wt / scale[:, None]
        Wd = dv.to_dev(Wp, torch.float32)
        cache.clear()
        model._dev['specW_keep'] = (F, cosmo)
        cache[key] = (Wd, split_tf32(Wd), dv.to_dev(scale, torch.float64))
    return cache[key]


def _spectra_contract(model, spectra, op):
    Wd, W_split, scale = op
    host = not dv.is_device_tensor(spectra)
    S = dv.to_dev(spectra, torch.float32)
    one = S.dim() == 1
    S = S.reshape(1, -1) if one else S
    nlp = Wd.shape[1]
    if S.shape[1] != nlp:
        S = torch.nn.functional.pad(S, (0, nlp - S.shape[1]))
    # spectra span many decades between galaxies: scale each row to unit maximum (a power of two: exact)
    rmax = S.abs().amax(dim=1).clamp_min(1e-30)
    rs = torch.pow(2.0, torch.ceil(torch.log2(rmax)))      # (torch.exp2 is a jiterator kernel: needs NVRTC at run time)
    out = gemm(S / rs[:, None], Wd, B_split=W_split).to(torch.float64) * rs.to(torch.float64)[:, None] * scale[None, :]
    out = out[0] if one else out
    return dv.to_host(out) if host else out


def spectra_Lnu(model, F, spectra):
    """core.sed.get_Lnu(F) (examples/SED_spectra.py:103) for one spectrum [n_lam] or a batch
    [G, n_lam] (e.g. generate_SED_batch(...)['total']): [F] / [G, F] rest-frame band
    luminosities, one tensor-core contraction spectra x filter weights."""
    return _spectra_contract(model, spectra, _filter_weight_operand(model, F, 'Lnu'))


def spectra_Fnu(model, F, spectra, z, cosmo, include_IGM=True):
    """core.sed.get_fnu(cosmo, z) followed by get_Fnu(F) (examples/SED_spectra.py:112): observed
    band fluxes in nJy, F defined on model.lam * (1 + z)."""
    return _spectra_contract(model, spectra, _filter_weight_operand(model, F, 'Fnu', z, cosmo, include_IGM))


# --------------------------------------------------------------------------
# emission lines
# --------------------------------------------------------------------------

class EmissionLines(_NGPGrid):
    """reference: synthobs/sed/models.py:326-448.  Same NGP look-up, LOS dust and sum over
    particles as generate_Lnu, over the `lines.p` grids: per (possibly blended, 'a,b') line the
    kernel carries two tables -- luminosity (transmitted part only: 10**luminosity, young
    particles) and continuum (incident / transmitted+nebular) -- with the dust curves
    evaluated at the mean line wavelength.  All requested lines share ONE so_sed_broadband
    pass (get_line_luminosities), where the reference loops over lines and particles."""

    def __init__(self, SPSIMF, dust_ISM=False, dust_BC=False, verbose=True,
                 path_to_SPS_grid=FLARE_dir + '/data/SPS/nebular/3.0/'):

        self.SPSIMF = SPSIMF

        if isinstance(SPSIMF, dict):   # convenience for synthetic grids: pass the dict itself
            self.grid = SPSIMF
        else:
            self.grid = pickle.load(open(f'{path_to_SPS_grid}/{SPSIMF}/lines.p', 'rb'), encoding='latin1')

        self.lines = self.grid['lines']

        self.lam = {l: self.grid[l]['lam'] for l in self.lines}
        self.lams = np.array([self.lam[l] for l in self.lines])

        if verbose:
            print('Available lines:')
            for lam, line in sorted(zip(self.lams, self.lines)):
                print(f'{line}')

        self.dust_BC = dust_BC
        self.dust_ISM = dust_ISM

        if self.dust_ISM:
            dust_model, dust_model_params = self.dust_ISM
            self.dust_curve_ISM = getattr(dust_curves, dust_model)(params=dust_model_params)

        if self.dust_BC:
            dust_model, dust_model_params = self.dust_BC
            self.dust_curve_BC = getattr(dust_curves, dust_model)(params=dust_model_params)

        self.units = {'luminosity': 'erg/s', 'nebular_continuum': 'erg/s/Hz', 'stellar_incident_continuum': 'erg/s/Hz',
                      'stellar_transmitted_continuum': 'erg/s/Hz', 'continuum': 'erg/s/Hz', 'EW': 'AA'}

        if 'stellar_transmitted_continuum' not in self.grid[self.lines[0]]:
            # canonical derivation (models.py:363-367)
            for l in self.lines:
                self.grid[l]['stellar_transmitted_continuum'] = self.grid[l]['stellar_continuum']
                self.grid[l]['stellar_incident_continuum'] = self.grid[l]['stellar_continuum']

        self._dev = {}

    def _line_tables(self, lines):
        """Two kernel tables per requested line: [luminosity, continuum] (see class docstring)."""
        key = tuple(lines)
        cache = self._dev.setdefault('tabs', {})
        if key not in cache:
            ncell = self.na * self.nz
            inc = np.zeros((2 * len(lines), ncell))
            tn = np.zeros((2 * len(lines), ncell))
            for i, line in enumerate(lines):
                for l in line.split(','):
                    g = self.grid[l]
                    tn[2 * i] += (10 ** np.asarray(g['luminosity'], dtype=np.float64)).reshape(-1)       # models.py:433
                    inc[2 * i + 1] += np.asarray(g['stellar_incident_continuum'], dtype=np.float64).reshape(-1)
                    tn[2 * i + 1] += (np.asarray(g['stellar_transmitted_continuum'], dtype=np.float64)
                                      + np.asarray(g['nebular_continuum'], dtype=np.float64)).reshape(-1)  # :434
            cache.clear()
            cache[key] = _scaled_tables(inc, tn)
        return cache[key]

    def _line_sums(self, lines, Masses, Ages, Metallicities, tauVs_ISM, tauVs_BC, fesc, log10t_BC, offsets=None):
        _cabi.require_device()
        lam = [np.mean([self.grid[l]['lam'] for l in line.split(',')]) for line in lines]              # models.py:401
        kI = [float(self.dust_curve_ISM.tau(w)) for w in lam for _ in (0, 1)] if self.dust_ISM else None
        kB = [float(self.dust_curve_BC.tau(w)) for w in lam for _ in (0, 1)] if self.dust_BC else None
        sums = []
        for c0 in range(0, len(lines), 16):       # so_sed_broadband takes up to 32 tables per call
            sub = lines[c0:c0 + 16]
            _, s_, _ = _broadband_launch(self, self._line_tables(sub), None if kI is None else kI[2 * c0:2 * c0 + 2 * len(sub)],
                                         None if kB is None else kB[2 * c0:2 * c0 + 2 * len(sub)],
                                         Masses, Ages, Metallicities, tauVs_ISM, tauVs_BC, fesc, log10t_BC,
                                         offsets=offsets)
            sums.append(s_)
        return torch.cat(sums, dim=1), lam

    def _finish(self, line, lam, luminosity, continuum):
        nl = float(len(line.split(',')))
        o = {'luminosity': luminosity, 'continuum': continuum}
        with np.errstate(divide='ignore', invalid='ignore'):
            total_continuum = (o['continuum'] / nl) * (3E8) / ((lam / nl) ** 2 * 1E-10)                # models.py:442
            o['EW'] = o['luminosity'] / total_continuum
        return o

    def get_line_luminosities(self, lines, Masses, Ages, Metallicities, tauVs_ISM=False, tauVs_BC=False, fesc=0.0,
                              log10t_BC=7., verbose=False):
        """reference: models.py:371-379 -- {line: {'luminosity', 'continuum', 'EW'}}; one kernel pass."""
        print('--- calculate quantities for multiple emission lines')
        print(lines)
        lines = list(lines)
        sums, lam = self._line_sums(lines, Masses, Ages, Metallicities, tauVs_ISM, tauVs_BC, fesc, log10t_BC)
        s = sums[0].cpu().numpy()
        o = {}
        for i, line in enumerate(lines):
            o[line] = self._finish(line, lam[i], s[2 * i], s[2 * i + 1])
            if verbose:
                self._report(line, lam[i], o[line])
        return o

    def get_line_luminosity(self, line, Masses, Ages, Metallicities, tauVs_ISM=False, tauVs_BC=False, fesc=0.0,
                            log10t_BC=7., verbose=False):
        """reference: models.py:382-448."""
        sums, lam = self._line_sums([line], Masses, Ages, Metallicities, tauVs_ISM, tauVs_BC, fesc, log10t_BC)
        s = sums[0].cpu().numpy()
        o = self._finish(line, lam[0], s[0], s[1])
        if verbose:
            self._report(line, lam[0], o)
        return o

    def get_line_luminosities_batch(self, lines, Masses, Ages, Metallicities, offsets, tauVs_ISM=False, tauVs_BC=False,
                                    fesc=0.0, log10t_BC=7.):
        """Many galaxies in one call (CSR `offsets`, no upstream equivalent): {line: {quantity: [G]}}."""
        lines = list(lines)
        sums, lam = self._line_sums(lines, Masses, Ages, Metallicities, tauVs_ISM, tauVs_BC, fesc, log10t_BC,
                                    offsets=offsets)
        s = sums.cpu().numpy()
        return {line: self._finish(line, lam[i], s[:, 2 * i], s[:, 2 * i + 1]) for i, line in enumerate(lines)}

    def _report(self, line, lam, o):
        print(f'----- {line.split(",")}')
        print(f'line wavelength/\\AA: {lam}')
        with np.errstate(divide='ignore'):
            for k, v in o.items():
                print(f'log10({k}/{self.units[k]}): {np.log10(v):.2f}')
