"""Synthetic inputs of the shapes SynthObs consumes.

The $FLARE data folder (BPASS grids, BlueTides particles, filter curves) is not
available offline, so every test, the smoke run and bench.py draw their inputs
from here.  Everything is float64 NumPy on the host, because that is what the
reference API receives (reference: synthobs/core.py:13-24 defines the particle
schema X, Y, MetSurfaceDensities, Ages [Myr], Metallicities, Masses [Msol]).

Nothing in this module touches the oracle; only `hashed_catalogue` uses torch (on whatever device it is
given), everything else is NumPy on the host.
"""

import os
import pickle

import numpy as np

# BPASS v2.2.1 metallicity nodes: not uniform in log Z
BPASS_Z = np.array([1e-5, 1e-4, 1e-3, 2e-3, 3e-3, 4e-3, 6e-3, 8e-3, 1e-2, 1.4e-2, 2e-2, 3e-2, 4e-2])

# particle mass used by the reference loaders (synthobs/core.py:22)
BLUETIDES_MASS = 1e10 * 5.90556119e-05 / 0.697


def ssp_grid(n_lam=9000, n_age=51, seed=0, with_incident=False):
    """BPASS-shape SSP grid dict with the keys define_model reads
    (reference: synthobs/sed/models.py:21-35)."""
    rng = np.random.default_rng(seed)
    log10age = 6.0 + 0.1 * np.arange(n_age)
    log10Z = np.log10(BPASS_Z)
    lam = np.logspace(1, 5, n_lam)
    nZ = log10Z.size
    amp = 10.0 ** (20.0 - 2.0 * (log10age[:, None] - 6.0) + 0.3 * rng.standard_normal((n_age, nZ)))
    shape = (lam / 1500.0) ** -0.5 * np.exp(-912.0 / lam)
    stellar = amp[:, :, None] * shape[None, None, :] * (1.0 + 0.05 * rng.standard_normal((n_age, nZ, n_lam)))
    stellar = np.abs(stellar)
    nebular = 0.2 * stellar * rng.random((n_age, nZ, n_lam))
    log10Q = 40.0 + rng.random((n_age, nZ))
    grid = {'lam': lam, 'log10age': log10age, 'log10Z': log10Z,
            'stellar': stellar, 'nebular': nebular, 'log10Q': log10Q}
    if with_incident:
        grid['stellar_incident'] = stellar.copy()
        tr = stellar.copy()
        tr[:, :, lam < 912] = 0.0
        grid['stellar_transmitted'] = tr
    return grid


def write_grid_pickle(grid, root, name='synthetic_grid'):
    """Write <root>/<name>/nebular.p so define_model(name, path_to_SPS_grid=root+'/')
    loads it (reference: synthobs/sed/models.py:19-21)."""
    d = os.path.join(root, name)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, 'nebular.p'), 'wb') as fh:
        pickle.dump(grid, fh, protocol=2)
    return name


LINE_LIST = {'HI3750': 3750.15, 'HI4861': 4861.32, 'HI6563': 6562.80, 'OIII4959': 4958.91, 'OIII5007': 5006.84,
             'OII3726': 3726.03, 'OII3729': 3728.81, 'NII6583': 6583.45, 'CIII1907': 1906.68, 'CIII1909': 1908.73}


def lines_grid(n_age=51, seed=0, with_incident=False):
    """`lines.p`-shape emission-line grid dict with the keys EmissionLines reads
    (reference: synthobs/sed/models.py:334-367,401,433-438): 'lines', 'log10age', 'log10Z'
    and per line 'lam', 'luminosity' (log10 erg/s per Msol), 'nebular_continuum',
    'stellar_continuum' (erg/s/Hz per Msol)."""
    rng = np.random.default_rng(seed + 77)
    log10age = 6.0 + 0.1 * np.arange(n_age)
    log10Z = np.log10(BPASS_Z)
    nZ = log10Z.size
    grid = {'lines': list(LINE_LIST), 'log10age': log10age, 'log10Z': log10Z}
    for name, lam in LINE_LIST.items():
        lum = 34.0 - 2.5 * (log10age[:, None] - 6.0) + 0.4 * rng.standard_normal((n_age, nZ))
        cont = 10.0 ** (20.0 - 2.0 * (log10age[:, None] - 6.0) + 0.3 * rng.standard_normal((n_age, nZ)))
        g = {'lam': lam, 'luminosity': lum, 'nebular_continuum': 0.2 * cont * rng.random((n_age, nZ)),
             'stellar_continuum': cont}
        if with_incident:
            g['stellar_incident_continuum'] = cont.copy()
            g['stellar_transmitted_continuum'] = cont * (0.5 + 0.5 * rng.random((n_age, nZ)))
        grid[name] = g
    return grid


def write_lines_pickle(grid, root, name='synthetic_grid'):
    """Write <root>/<name>/lines.p so EmissionLines(name, path_to_SPS_grid=root) loads it
    (reference: synthobs/sed/models.py:334)."""
    d = os.path.join(root, name)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, 'lines.p'), 'wb') as fh:
        pickle.dump(grid, fh, protocol=2)
    return name


class Filter:
    """Stand-in for a flare.filters filter object: .lam, .T, .pivwv()
    (protocol used at synthobs/sed/models.py:43,47,97)."""

    def __init__(self, lam, T):
        self.lam = np.asarray(lam, dtype=np.float64)
        self.T = np.asarray(T, dtype=np.float64)

    def pivwv(self):
        return np.sqrt(np.trapezoid(self.lam * self.T, x=self.lam) / np.trapezoid(self.T / self.lam, x=self.lam))


def filter_set(lam, n_filters=10, kind='tophat', lo=1200.0, hi=30000.0, prefix='FAKE.TH.'):
    """n broadband filters on the wavelength grid `lam` (the caller passes
    model.lam for rest-frame filters or model.lam*(1+z) for observed ones, as
    flare.filters.add_filters(new_lam=...) does). Returns the F dict:
    F['filters'] list + F[name] objects."""
    lam = np.asarray(lam, dtype=np.float64)
    centres = np.logspace(np.log10(lo), np.log10(hi), n_filters)
    F = {'filters': []}
    for i, c in enumerate(centres):
        name = '%s%02d' % (prefix, i)
        if kind == 'tophat':
            T = ((lam > c * 0.85) & (lam < c * 1.15)).astype(np.float64)
        else:
            T = np.exp(-0.5 * ((lam - c) / (0.08 * c)) ** 2)
            T[T < 1e-8] = 0.0
        F['filters'].append(name)
        F[name] = Filter(lam, T)
    return F


class _Quantity:
    def __init__(self, v):
        self.value = v

    def to(self, unit):
        return self


class FixedCosmology:
    """Cosmology provider with the two methods the hot path calls
    (synthobs/sed/models.py:59, synthobs/morph/images.py:137). Values are for
    z=8 in a Planck-like cosmology, stated here so runs are reproducible."""

    def __init__(self, dl_cm=2.53e29, arcsec_per_kpc=0.2033):
        self.dl_cm = dl_cm
        self.arcsec_per_kpc = arcsec_per_kpc

    def luminosity_distance(self, z):
        return _Quantity(self.dl_cm)

    def arcsec_per_kpc_proper(self, z):
        return _Quantity(self.arcsec_per_kpc)


class GaussianPSF:
    """Analytic Gaussian PSF provider, same protocol and formula as the
    reference's SpitzerPSF (synthobs/morph/PSF.py:139-154): .f(x1d, y1d) ->
    [ny, nx] with x, y in native pixels."""

    def __init__(self, fwhm_pix):
        self.FWHM = float(fwhm_pix)

    def f(self, x, y):
        xx, yy = np.meshgrid(x, y)
        return np.exp(-4 * np.log(2) * (xx ** 2 + yy ** 2) / self.FWHM ** 2)


class TabulatedPSF:
    """Tabulated PSF provider with the attributes the reference's FITS-backed PSFs carry
    (synthobs/morph/PSF.py:69-105 HubblePSF: .data, .ndim, .width ["], .f = bilinear
    interpolation over native pixels with 0 outside) -- the protocol `point()` needs
    (synthobs/morph/images.py:384-392).  `sub` = samples per native pixel."""

    def __init__(self, data, sub, pixel_scale):
        from scipy.interpolate import RegularGridInterpolator
        self.data = np.asarray(data, dtype=np.float64)
        self.ndim = self.data.shape[0]
        self.width = self.ndim * pixel_scale / sub
        x = np.linspace(-(self.ndim / 2.) / sub, (self.ndim / 2.) / sub, self.ndim)
        self._interp = RegularGridInterpolator((x, x), self.data, method='linear', bounds_error=False, fill_value=0.0)

    def f(self, x, y):
        yy, xx = np.meshgrid(y, x, indexing='ij')
        return self._interp(np.stack([yy.ravel(), xx.ravel()], axis=1)).reshape(len(y), len(x))


def airy_like_psf(ndim=101, sub=5, fwhm_pix=2.0, pixel_scale=0.031, seed=0):
    """A non-separable, slightly asymmetric tabulated PSF (core + ring + diffraction-spike-like
    cross) so the tabulated-PSF path is not exercised by a pure Gaussian only."""
    rng = np.random.default_rng(seed)
    x = np.linspace(-(ndim / 2.) / sub, (ndim / 2.) / sub, ndim)
    xx, yy = np.meshgrid(x, x)
    r = np.hypot(xx, yy)
    core = np.exp(-4 * np.log(2) * r ** 2 / fwhm_pix ** 2)
    ring = 0.05 * np.exp(-0.5 * ((r - 1.6 * fwhm_pix) / (0.3 * fwhm_pix)) ** 2)
    spikes = 0.02 * (np.exp(-np.abs(xx) / 0.2) + np.exp(-np.abs(yy + 0.1 * xx) / 0.2)) * np.exp(-r / (3 * fwhm_pix))
    data = (core + ring + spikes) * (1.0 + 0.01 * rng.standard_normal((ndim, ndim)))
    return TabulatedPSF(np.abs(data), sub, pixel_scale)


def particles(n, seed=1, lognormal_mass=False, clumps=3, scale_kpc=1.5):
    """Star-particle catalogue: clustered positions (wide kNN range), ages in
    Myr, metallicities, masses, LOS metal surface densities and the two optical
    depths built exactly as the examples do (examples/SED_luminosities.py:69-70)."""
    rng = np.random.default_rng(seed)
    X = rng.normal(0.0, scale_kpc, n)
    Y = rng.normal(0.0, scale_kpc, n)
    for c in range(clumps):
        m = int(0.05 * n)
        if m == 0:
            break
        idx = rng.choice(n, m, replace=False)
        cx, cy = rng.normal(0, scale_kpc, 2)
        X[idx] = cx + rng.normal(0, 0.05 * scale_kpc, m)
        Y[idx] = cy + rng.normal(0, 0.05 * scale_kpc, m)
    Ages = 10.0 ** rng.uniform(0.0, 4.2, n)
    Z = 10.0 ** rng.uniform(-4.5, -1.5, n)
    if lognormal_mass:
        M = BLUETIDES_MASS * np.exp(0.3 * rng.standard_normal(n))
    else:
        M = np.ones(n) * BLUETIDES_MASS
    msd = np.exp(rng.normal(np.log(0.3 / 10 ** 5.2), 0.8, n))
    out = {'X': X, 'Y': Y, 'Ages': Ages, 'Metallicities': Z, 'Masses': M,
           'MetSurfaceDensities': msd,
           'tauVs_ISM': (10 ** 5.2) * msd, 'tauVs_BC': 2.0 * (Z / 0.01)}
    return out


def galaxy_batch(n_galaxies, n_min=1000, n_max=100000, seed=2, fixed_n=None):
    """Concatenated catalogue of many galaxies + CSR offsets.
    Sizes are 10^U(log n_min, log n_max) unless fixed_n is given."""
    rng = np.random.default_rng(seed)
    if fixed_n is not None:
        sizes = np.full(n_galaxies, int(fixed_n), dtype=np.int64)
    else:
        sizes = (10.0 ** rng.uniform(np.log10(n_min), np.log10(n_max), n_galaxies)).astype(np.int64)
    offsets = np.zeros(n_galaxies + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    total = int(offsets[-1])
    p = particles(total, seed=seed + 1000, clumps=0)
    p['offsets'] = offsets
    return p


def adversarial_ages(log10age_grid, ulps=(0, 1, 2)):
    """Ages [Myr] placed at the midpoints between age nodes +- a few ulp, the
    values where a device log10 one ulp off glibc's would flip the NGP index
    (SURVEY.md H3)."""
    mids = 0.5 * (log10age_grid[:-1] + log10age_grid[1:])
    base = 10.0 ** (mids - 6.0)
    out = [base]
    for u in ulps:
        if u == 0:
            continue
        up, dn = base.copy(), base.copy()
        for _ in range(u):
            up = np.nextafter(up, np.inf)
            dn = np.nextafter(dn, -np.inf)
        out += [up, dn]
    return np.concatenate(out)


# ---------------------------------------------------------------------------
# counter-based catalogue: every particle's attributes are a pure function of its GLOBAL id, so any rank can
# generate exactly its own share of ONE fixed catalogue (bench.py config 4: 1e4 galaxies sharded over the GPUs)
# ---------------------------------------------------------------------------

def _wrap64(c):
    c &= (1 << 64) - 1
    return c - (1 << 64) if c >= (1 << 63) else c


def _lsr(z, k):
    return (z >> k) & ((1 << (64 - k)) - 1)


def hashed_uniform(ids, stream):
    """U(0,1) float64 per element of the int64 tensor `ids` (splitmix64 of id and stream number; torch int64
    arithmetic wraps like uint64)."""
    import torch
    z = ids * _wrap64(0x9E3779B97F4A7C15) + _wrap64((stream + 1) * 0xD1B54A32D192ED03)
    z = (z ^ _lsr(z, 30)) * _wrap64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _wrap64(0x94D049BB133111EB)
    z = z ^ _lsr(z, 31)
    return (_lsr(z, 11).to(torch.float64) + 0.5) * (1.0 / 9007199254740992.0)


def hashed_catalogue(ids, scale_kpc=1.5):
    """Particle attributes for the global particle ids `ids` (int64 tensor, any device): the same distributions
    as particles(..., clumps=0) -- Gaussian positions, Ages = 10^U(0, 4.2) Myr, Z = 10^U(-4.5, -1.5), constant
    mass, log-normal metal surface density, tauVs as in examples/SED_luminosities.py:69-70."""
    import math
    import torch
    u = [hashed_uniform(ids, k) for k in range(7)]
    r = torch.sqrt(-2.0 * torch.log(u[0]))
    X = scale_kpc * r * torch.cos(2.0 * math.pi * u[1])
    Y = scale_kpc * r * torch.sin(2.0 * math.pi * u[1])
    Ages = torch.pow(10.0, 4.2 * u[2])
    Z = torch.pow(10.0, -4.5 + 3.0 * u[3])
    g = torch.sqrt(-2.0 * torch.log(u[4])) * torch.cos(2.0 * math.pi * u[5])
    msd = torch.exp(math.log(0.3 / 10 ** 5.2) + 0.8 * g)
    M = torch.full_like(X, BLUETIDES_MASS)
    return {'X': X, 'Y': Y, 'Ages': Ages, 'Metallicities': Z, 'Masses': M, 'MetSurfaceDensities': msd,
            'tauVs_ISM': (10 ** 5.2) * msd, 'tauVs_BC': 2.0 * (Z / 0.01), 'weight': 0.5 + u[6]}


def ragged_sizes(n_galaxies, n_min=1000, n_max=100000, seed=2):
    """galaxy sizes 10^U(log n_min, log n_max) and their CSR offsets"""
    rng = np.random.default_rng(seed)
    sizes = (10.0 ** rng.uniform(np.log10(n_min), np.log10(n_max), n_galaxies)).astype(np.int64)
    offsets = np.zeros(n_galaxies + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    return sizes, offsets
