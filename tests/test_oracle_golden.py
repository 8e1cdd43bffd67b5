"""CPU: the oracle (oracle/*.py) against the golden vectors produced by the
unmodified reference.  These pin the oracle; the -m gpu tests then compare the
CUDA path with both."""

import numpy as np
import pytest

from oracle import image_oracle, sed_oracle
from synthobs_b200 import providers, synthetic

import _cases

RT = 1e-12


@pytest.mark.parametrize('tag', _cases.SED_CASES)
def test_sed_oracle_matches_reference(tag):
    g, grid, F, dust_ISM, dust_BC, kw = _cases.sed_inputs(tag)
    sed_oracle.ensure_incident(grid)
    z = float(g['z'])
    observed = bool(g['observed'])
    if observed:
        tabs = sed_oracle.fnu_grid(grid, F, z, float(g['dl_cm']), providers.madau)
    else:
        tabs = sed_oracle.lnu_grid(grid, F)
    args = (grid['log10age'], grid['log10Z'], g['Masses'], g['Ages'], g['Metallicities'])
    pkw = dict(fesc=kw['fesc'], log10t_BC=kw['log10t_BC'], tauVs_ISM=kw.get('tauVs_ISM'), tauVs_BC=kw.get('tauVs_BC'))
    kI, kB = {}, {}
    for fi, f in enumerate(F['filters']):
        for c in sed_oracle.COMPONENTS:
            np.testing.assert_allclose(tabs[f][c], g['tab_%d_%s' % (fi, c)], rtol=RT)
        piv = F[f].pivwv() / (1. + z) if observed else F[f].pivwv()
        kI[f] = _cases.dust_k(dust_ISM, piv)
        kB[f] = _cases.dust_k(dust_BC, piv)
        arr = sed_oracle.broadband_array(tabs[f], *args, k_ISM=kI[f], k_BC=kB[f], **pkw)
        np.testing.assert_allclose(arr, g['arr_%d' % fi], rtol=RT)
    sums = sed_oracle.broadband(tabs, F['filters'], *args, k_ISM=kI if dust_ISM else None,
                                k_BC=kB if dust_BC else None, **pkw)
    np.testing.assert_allclose([sums[f] for f in F['filters']], g['sums'], rtol=1e-11)
    np.testing.assert_allclose(sed_oracle.log10Q(grid, g['Masses'], g['Ages'], g['Metallicities']), g['log10Q'], rtol=RT)
    o = sed_oracle.sed_spectra(grid, g['Masses'], g['Ages'], g['Metallicities'], tauVs_ISM=kw.get('tauVs_ISM'),
                               tauVs_BC=kw.get('tauVs_BC'), k_ISM_lam=_cases.dust_k(dust_ISM, grid['lam']),
                               k_BC_lam=_cases.dust_k(dust_BC, grid['lam']), fesc=kw['fesc'], log10t_BC=kw['log10t_BC'])
    for k in ('stellar', 'intrinsic', 'BC', 'total', 'no_HII_no_BC'):
        np.testing.assert_allclose(o[k], g['sed_' + k], rtol=1e-11, atol=0)
    for k in ('f_esc', 'tau', 'tau_relV'):
        np.testing.assert_allclose(o[k], g['sed_' + k], rtol=1e-8, atol=1e-12, equal_nan=True)


@pytest.mark.parametrize('tag', _cases.LINES_CASES)
def test_line_oracle_matches_reference(tag):
    """oracle line_luminosity vs EmissionLines.get_line_luminosity[ies] of the unmodified reference"""
    g, lgrid, dust_ISM, dust_BC, kw = _cases.lines_inputs(tag)
    sed_oracle.ensure_line_continua(lgrid)
    pkw = dict(fesc=kw['fesc'], log10t_BC=kw['log10t_BC'], tauVs_ISM=kw.get('tauVs_ISM'), tauVs_BC=kw.get('tauVs_BC'))

    def one(line):
        lam = np.mean([lgrid[l]['lam'] for l in line.split(',')])
        return sed_oracle.line_luminosity(lgrid, line, g['Masses'], g['Ages'], g['Metallicities'],
                                          k_ISM=_cases.dust_k(dust_ISM, lam), k_BC=_cases.dust_k(dust_BC, lam), **pkw)
    for i, line in enumerate(g['lines']):
        o = one(str(line))
        for q in ('luminosity', 'continuum', 'EW'):
            np.testing.assert_allclose(o[q], g[q][i], rtol=1e-11)
    o = one('NII6583')
    for q in ('luminosity', 'continuum', 'EW'):
        np.testing.assert_allclose(o[q], g['single_' + q], rtol=1e-11)


def test_sed_indices_match_reference():
    g = _cases.load('sed_indices')
    grid = synthetic.ssp_grid(**_cases.GRID_KW)
    ia, iZ, log10age = sed_oracle.ngp_indices(grid['log10age'], grid['log10Z'], g['Ages'], g['Metallicities'])
    assert np.array_equal(ia, g['ia'])
    assert np.array_equal(iZ, g['iZ'])
    with np.errstate(invalid='ignore'):
        assert np.array_equal(log10age < 7.0, g['young'])


def test_histogram_gemm_formulation_equals_loop():
    """SURVEY §8a mapping: dust-free spectra == cell histogram contracted with the grid."""
    g, grid, F, dI, dB, kw = _cases.sed_inputs('lnu_nodust')
    sed_oracle.ensure_incident(grid)
    H = sed_oracle.cell_histogram(grid['log10age'], grid['log10Z'], g['Masses'], g['Ages'], g['Metallicities'])
    nl = grid['lam'].size
    inc = grid['stellar_incident'].reshape(-1, nl)
    stellar = (H[0] + H[1]).reshape(-1) @ inc
    np.testing.assert_allclose(stellar, g['sed_stellar'], rtol=1e-11)


@pytest.mark.parametrize('tag', _cases.IMG_CASES)
def test_image_oracle_matches_reference(tag):
    g = _cases.load('img_' + tag)
    sm = _cases.smoothing_of(g)
    adaptive = bool(sm) and sm[0] == 'adaptive'
    # literal per-particle loop, separable form, and the windowed/tiled form the full-size GPU parity tests use
    for literal, tiled in ([(True, False), (False, False), (False, True)] if adaptive else [(False, False)]):
        with np.errstate(all='ignore'):
            o = image_oracle.core(g['X'], g['Y'], g['L'], resolution=float(g['resolution']), ndim=int(g['ndim']),
                                  smoothing=sm, literal=literal, tiled=tiled)
        assert np.array_equal(o['hist'], g['hist'])
        np.testing.assert_allclose(o['simple'], g['simple'], rtol=RT)
        peak = np.abs(g['data']).max()
        np.testing.assert_allclose(o['data'], g['data'], rtol=1e-9, atol=1e-13 * peak)


def test_windowed_oracle_equals_separable():
    """adaptive_window / adaptive_tiled (terms that are exactly 0.0 in float64 skipped) == adaptive_separable, incl.
    multi-channel weights, a window of a larger image, dk = inf (k > N) and the FWHM floor."""
    rng = np.random.default_rng(5)
    for n, ndim, res, k in ((4000, 96, 0.1, 8), (9000, 200, 0.02, 5), (20, 40, 0.1, 64)):
        p = synthetic.particles(n, seed=n)
        width, G = image_oracle.pixel_grid(res, ndim)
        sel = image_oracle.select(p['X'], p['Y'], width)
        X, Y = p['X'][sel], p['Y'][sel]
        L = np.stack([p['Masses'][sel] * rng.random(X.size), rng.random(X.size)])
        dk = image_oracle.kth_neighbour_distance(X, Y, k)
        dk[::7] = 0.0          # FWHM floor (images.py:89)
        full = np.stack([image_oracle.adaptive_separable(G, X, Y, L[c], dk) for c in range(2)])
        tiled = image_oracle.adaptive_tiled(G, X, Y, L, dk, block=37)
        win = image_oracle.adaptive_window(G, X, Y, L, dk, (5, 33), (11, 40))
        for c in range(2):
            peak = full[c].max()
            np.testing.assert_allclose(tiled[c], full[c], rtol=1e-12, atol=1e-15 * peak)
            np.testing.assert_allclose(win[c], full[c][5:33, 11:40], rtol=1e-12, atol=1e-15 * peak)


def test_pixel_indices_match_reference():
    g = _cases.load('img_indices')
    width, G = image_oracle.pixel_grid(float(g['resolution']), int(g['ndim']))
    assert np.array_equal(image_oracle.ngp_pixels(G, g['x']), g['idx'])


def test_observed_oracle_matches_reference():
    g = _cases.load('img_observed')
    fwhm = g['fwhm']
    native = 0.031
    geo = image_oracle.observed_geometry(native, float(g['arcsec_per_kpc']), float(g['width_arcsec']),
                                         super_sampling=int(g['super_sampling']), xoffset_pix=0.25, yoffset_pix=-0.4)
    o = image_oracle.observed_particle(geo, g['X'], g['Y'], g['L_0'], synthetic.GaussianPSF(fwhm[0]).f,
                                       smoothing=('adaptive', 8.))
    np.testing.assert_allclose(o['X'], g['obs_X_after'], rtol=0, atol=0)
    for a, b in (('super_no_PSF', 'obs_super_no_PSF'), ('super_data', 'obs_super_data'), ('no_PSF', 'obs_no_PSF'),
                 ('data', 'obs_data')):
        peak = np.abs(g[b]).max()
        np.testing.assert_allclose(o[a], g[b], rtol=1e-9, atol=1e-12 * peak)
    assert geo['pixel_scale'] == float(g['obs_pixel_scale'])
    assert geo['pixel_scale_kpc'] == float(g['obs_pixel_scale_kpc'])


def test_flux_conservation_property():
    """images.py:95-97: every adaptive Gaussian is normalised to its particle's l."""
    g = _cases.load('img_adaptive16')
    o = image_oracle.core(g['X'], g['Y'], g['L'], resolution=float(g['resolution']), ndim=int(g['ndim']),
                          smoothing=('adaptive', 16), return_aux=True)
    np.testing.assert_allclose(o['data'].sum(), g['L'][o['aux']['sel']].sum(), rtol=1e-12)
    assert o['hist'].sum() == o['aux']['sel'].sum()
    assert image_oracle.rebin(o['data'], (25, 25)).sum() == pytest.approx(o['data'].sum(), rel=1e-13)


def test_analytic_sources_match_reference():
    """oracle observed_sersic / point_source vs observed.Sersic, Sersic(), point() of the unmodified reference"""
    g = _cases.load('img_analytic')
    cosmo = synthetic.FixedCosmology()
    p = dict(zip(('r_eff', 'n', 'ellip', 'theta'), g['sersic_p']))
    geo = image_oracle.observed_geometry(0.031, cosmo.arcsec_per_kpc, 0.8, resampling_factor=2, super_sampling=3,
                                         xoffset_pix=0.3, yoffset_pix=-0.2)
    o = image_oracle.observed_sersic(geo, 3.5e28, p, synthetic.GaussianPSF(2.1).f, providers.Sersic2D)
    for k in ('super_no_PSF', 'super_data', 'no_PSF', 'data'):
        np.testing.assert_allclose(o[k], g['ser_' + k], rtol=1e-9, atol=1e-12 * np.abs(g['ser_' + k]).max())
    assert geo['pixel_scale'] == g['ser_pixel_scale'] and geo['pixel_scale_kpc'] == g['ser_pixel_scale_kpc']
    for fi, (ps, fw, L) in enumerate(((0.031, 2.1, 1.0e28), (0.063, 1.8, 2.5e28))):
        geo = image_oracle.observed_geometry(ps, cosmo.arcsec_per_kpc, 1.2, super_sampling=2)
        o = image_oracle.observed_sersic(geo, L, {'r_eff': 0.5, 'n': 4.0, 'ellip': 0.0, 'theta': 0.0},
                                         synthetic.GaussianPSF(fw).f, providers.Sersic2D)
        np.testing.assert_allclose(o['data'], g['sermulti_%d_data' % fi], rtol=1e-9, atol=1e-12 * o['data'].max())
        np.testing.assert_allclose(o['no_PSF'], g['sermulti_%d_no_PSF' % fi], rtol=1e-9)
    psf = synthetic.airy_like_psf(ndim=101, sub=5, fwhm_pix=2.0, pixel_scale=0.031)
    o = image_oracle.point_source(7.0e3, 0.031, 0.55, psf, super_sampling=5, xoffset_pix=0.21, yoffset_pix=-0.37)
    assert np.array_equal(o['super_hist'], g['pt_super_hist']) and np.array_equal(o['super_simple'], g['pt_super_simple'])
    for k in ('super_data', 'no_PSF', 'data'):
        np.testing.assert_allclose(o[k], g['pt_' + k], rtol=1e-9, atol=1e-12 * np.abs(g['pt_' + k]).max())
    assert o['ndim'] == int(g['pt_ndim']) and o['pixel_scale'] == float(g['pt_pixel_scale'])
    psf2 = synthetic.airy_like_psf(ndim=75, sub=5, fwhm_pix=1.4, pixel_scale=0.063, seed=4)
    o = image_oracle.point_source(2.0, 0.063, 1.5, psf2, resampling_factor=2, super_sampling=3)
    for k in ('super_data', 'no_PSF', 'data'):
        np.testing.assert_allclose(o[k], g['pt2_' + k], rtol=1e-9, atol=1e-12 * np.abs(g['pt2_' + k]).max())
