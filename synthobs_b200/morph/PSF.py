"""synthobs.morph.PSF — PSF providers (reference: synthobs/morph/PSF.py).

The hot path only needs the provider PROTOCOL: an object with
.f(x1d, y1d) -> [ny, nx] array, x/y in native pixels (used at
synthobs/morph/images.py:198-200).  The analytic providers are implemented; the
FITS-backed ones (WebbPSF, TinyTim/Hubble, Euclid) need $FLARE data files and
scipy.interpolate.interp2d, which scipy >= 1.14 removed — out of scope here and
they say so loudly.
"""

import numpy as np


class SpitzerPSF():
    """reference: PSF.py:139-154 — Gaussian with the IRAC FWHM, in native pixels."""

    def __init__(self, f):
        self.filter = f
        if f.split('.')[-1] == 'ch1':
            self.FWHM = 1.66
        if f.split('.')[-1] == 'ch2':
            self.FWHM = 1.72
        self.FWHM /= 1.22

    def f(self, x, y):
        xx, yy = np.meshgrid(x, y)
        return np.exp(-4 * np.log(2) * (xx ** 2 + yy ** 2) / self.FWHM ** 2)


def Spitzer(filters):
    return {f: SpitzerPSF(f) for f in filters}


class gauss():
    """reference: PSF.py:165-171 — unit-amplitude circular Gaussian of given FWHM
    (native pixels).  The reference wraps astropy's Gaussian2D, whose call
    signature is f(xx, yy) on meshgrids; this provider follows the .f(x1d, y1d)
    protocol the imaging code actually uses."""

    def __init__(self, FWHM):
        self.FWHM = FWHM
        self.stddev = FWHM / (2 * np.sqrt(2 * np.log(2)))

    def f(self, x, y):
        xx, yy = np.meshgrid(x, y)
        return np.exp(-(xx ** 2 + yy ** 2) / (2.0 * self.stddev ** 2))


def _needs_data(name):
    class _Unavailable():
        def __init__(self, *a, **k):
            raise NotImplementedError(
                '%s needs FITS PSF files under $FLARE/data/PSF and scipy.interpolate.interp2d (removed in '
                'scipy>=1.14); not part of the B200 hot path. Pass any object with .f(x, y) instead.' % name)
    _Unavailable.__name__ = name
    return _Unavailable


WebbPSF = _needs_data('WebbPSF')
HubblePSF = _needs_data('HubblePSF')
EuclidPSF = _needs_data('EuclidPSF')


def PSF(f, **kwargs):
    """reference: PSF.py:26-35 — dispatch on the observatory prefix of the filter name
    (HST / Hubble, JWST / Webb, Euclid, Spitzer)."""
    obs = f.split('.')[0]
    if obs == 'Spitzer':
        return SpitzerPSF(f, **kwargs)
    if obs in ('JWST', 'Webb'):
        return WebbPSF(f, **kwargs)
    if obs in ('HST', 'Hubble'):
        return HubblePSF(f, **kwargs)
    if obs == 'Euclid':
        return EuclidPSF(f, **kwargs)
    raise KeyError(f)       # upstream: UnboundLocalError on `psf`


def PSFs(filters, **kwargs):
    return {f: PSF(f, **kwargs) for f in filters}
