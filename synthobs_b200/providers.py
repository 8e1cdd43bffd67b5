"""Offline host-side providers for the third-party protocols the hot path consumes.

The reference gets these from packages that are not vendored under
/root/reference and are not installed here: `interrogator.sed`
(dust_curves, IGM, core.sed) and `flare`.  Every quantity they produce is a
HOST-computed opaque input to the GPU path (dust k(lambda), IGM transmission,
filter weights), so parity against the oracle never depends on how faithful
these stand-ins are to upstream: oracle and CUDA path consume the same numbers.

If the real `interrogator` is importable it is used instead (drop-in behaviour);
otherwise these are.  Call sites mirrored: synthobs/sed/models.py:11-13,97,101,
163,167,240-244,257,262.
"""

import types

import numpy as np


class simple:
    """Power-law attenuation curve tau(lam)/tau_V = (lam/5500 A)^slope
    [recall of interrogator.sed.dust_curves.simple; used as
    getattr(dust_curves, name)(params=...).tau(lam) at models.py:97]."""

    def __init__(self, params={'slope': -1.0}):
        self.params = params

    def tau(self, lam):
        return (np.asarray(lam, dtype=np.float64) / 5500.0) ** self.params['slope']


class Calzetti_like:
    """A second, non-power-law curve so tests exercise distinct ISM/BC curves."""

    def __init__(self, params={}):
        self.params = params

    def tau(self, lam):
        x = 1e4 / np.asarray(lam, dtype=np.float64)  # 1/micron
        k = 2.659 * (-2.156 + 1.509 * x - 0.198 * x ** 2 + 0.011 * x ** 3) + 4.05
        k5500 = 2.659 * (-2.156 + 1.509 / 0.55 - 0.198 / 0.55 ** 2 + 0.011 / 0.55 ** 3) + 4.05
        return np.clip(k / k5500, 0.0, None)


def madau(lam_obs, z):
    """Madau (1995)-style mean IGM transmission on an observed-frame wavelength
    grid [A]: Lyman-series line blanketing + a smooth continuum term
    [restated from the published prescription; call site models.py:65-67]."""
    lam_obs = np.asarray(lam_obs, dtype=np.float64)
    lines = [1216.0, 1026.0, 972.5, 950.0]
    coeff = [0.0036, 0.0017, 0.0012, 0.00093]
    tau = np.zeros_like(lam_obs)
    for l, a in zip(lines, coeff):
        sel = lam_obs < l * (1.0 + z)
        tau[sel] += a * (lam_obs[sel] / l) ** 3.46
    sel = lam_obs < 912.0 * (1.0 + z)
    xc = lam_obs[sel] / 912.0
    xem = 1.0 + z
    tau[sel] += (0.25 * xc ** 3 * (xem ** 0.46 - xc ** 0.46)
                 + 9.4 * xc ** 1.5 * (xem ** 0.18 - xc ** 0.18)
                 - 0.7 * xc ** 3 * (xc ** -1.32 - xem ** -1.32)
                 - 0.023 * (xem ** 1.68 - xc ** 1.68))
    return np.exp(-np.clip(tau, 0.0, 700.0))


class sed:
    """Spectrum container with the attributes generate_SED fills
    (models.py:240-244: core.sed(lam, description=...) -> .lam, .lnu) and the
    post-processing the spectra examples call next (examples/SED_spectra.py:
    103,112): get_Lnu(F), get_fnu(cosmo, z), get_Fnu(F)."""

    def __init__(self, lam, description=False):
        self.description = description
        self.lam = lam
        self.lnu = np.zeros(np.shape(lam))

    def get_Lnu(self, F):
        self.Lnu = {f: np.trapezoid(self.lnu * F[f].T, x=F[f].lam) / np.trapezoid(F[f].T, x=F[f].lam)
                    for f in F['filters']}
        return self.Lnu

    def get_fnu(self, cosmo, z, include_IGM=True):
        self.lamz = self.lam * (1.0 + z)
        dl = cosmo.luminosity_distance(z).to('cm').value
        self.fnu = 1e23 * 1e9 * self.lnu * (1.0 + z) / (4.0 * np.pi * dl ** 2)
        if include_IGM:
            self.fnu = self.fnu * madau(self.lamz, z)
        return self.fnu

    def get_Fnu(self, F):
        self.Fnu = {f: np.trapezoid(self.fnu * F[f].T, x=F[f].lam) / np.trapezoid(F[f].T, x=F[f].lam)
                    for f in F['filters']}
        return self.Fnu

    def return_log10Q(self):
        """log10 of the ionising photon rate [1/s]: integral of L_nu / (h nu) d nu blueward of
        912 A = integral of lnu / (h lam) d lam  [recall of interrogator core.sed.return_log10Q;
        call site examples/SED_Lion.py:38]."""
        h = 6.626E-27  # erg s
        sel = self.lam < 912.
        return np.log10(np.trapezoid(self.lnu[sel] / (h * self.lam[sel]), x=self.lam[sel]))


class Sersic2D:
    """astropy.modeling.models.Sersic2D restated [recall of the published evaluate(); used at
    synthobs/morph/images.py:229-231 as Sersic2D(amplitude, r_eff, n, x_0, y_0, ellip, theta)(xx, yy)]:
    amplitude * exp(-b_n ((z)^(1/n) - 1)), z = sqrt((x_maj/a)^2 + (x_min/b)^2), a = r_eff,
    b = (1 - ellip) r_eff, b_n = gammaincinv(2n, 0.5)."""

    def __init__(self, amplitude=1, r_eff=1, n=4, x_0=0, y_0=0, ellip=0, theta=0):
        self.amplitude, self.r_eff, self.n = amplitude, r_eff, n
        self.x_0, self.y_0, self.ellip, self.theta = x_0, y_0, ellip, theta

    @staticmethod
    def bn(n):
        from scipy.special import gammaincinv
        return float(gammaincinv(2.0 * n, 0.5))

    def __call__(self, x, y):
        bn = self.bn(self.n)
        a, b = self.r_eff, (1 - self.ellip) * self.r_eff
        cos_theta, sin_theta = np.cos(self.theta), np.sin(self.theta)
        x_maj = (x - self.x_0) * cos_theta + (y - self.y_0) * sin_theta
        x_min = -(x - self.x_0) * sin_theta + (y - self.y_0) * cos_theta
        z = np.sqrt((x_maj / a) ** 2 + (x_min / b) ** 2)
        return self.amplitude * np.exp(-bn * (z ** (1 / self.n) - 1))


dust_curves = types.SimpleNamespace(simple=simple, Calzetti_like=Calzetti_like)
IGM = types.SimpleNamespace(madau=madau)
core = types.SimpleNamespace(sed=sed)


def resolve():
    """Return (core, IGM, dust_curves): upstream interrogator if installed,
    else the offline providers above."""
    try:
        from interrogator.sed import core as c, IGM as i, dust_curves as d  # noqa
        return c, i, d
    except Exception:
        return core, IGM, dust_curves
