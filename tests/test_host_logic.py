"""CPU: host-side logic of the API mirror (thresholds, filter weights, geometry)."""

import numpy as np
import pytest

from oracle import image_oracle, sed_oracle
from synthobs_b200 import synthetic
from synthobs_b200.sed import models

import _cases


def _count(thr, v):
    r = (v[:, None] >= thr[None, :]).sum(1)
    r[~(v < np.inf)] = 0
    return r


def test_thresholds_reproduce_reference_indices():
    g = _cases.load('sed_indices')
    grid = synthetic.ssp_grid(**_cases.GRID_KW)
    ta = models.ngp_thresholds(grid['log10age'], 6.0)
    tz = models.ngp_thresholds(grid['log10Z'], 0.0)
    assert np.array_equal(_count(ta, g['Ages']), g['ia'])
    assert np.array_equal(_count(tz, g['Metallicities']), g['iZ'])
    yt = models.young_threshold(7.0)
    assert np.array_equal((g['Ages'] < yt) & (g['Ages'] >= 0), g['young'])


def test_thresholds_random_dense():
    grid = synthetic.ssp_grid(n_lam=8)
    rng = np.random.default_rng(0)
    A = 10 ** rng.uniform(-2, 6, 200000)
    Z = 10 ** rng.uniform(-7, 0, 200000)
    ia, iZ, la = sed_oracle.ngp_indices(grid['log10age'], grid['log10Z'], A, Z)
    assert np.array_equal(_count(models.ngp_thresholds(grid['log10age'], 6.0), A), ia)
    assert np.array_equal(_count(models.ngp_thresholds(grid['log10Z'], 0.0), Z), iZ)
    for t in (6.5, 7.0, 7.3, 8.0):
        yt = models.young_threshold(t)
        assert np.array_equal(A < yt, la < t)


def test_trapz_weights_identity():
    rng = np.random.default_rng(1)
    x = np.sort(rng.uniform(0, 10, 200))
    y = rng.random(200)
    assert np.trapezoid(y, x=x) == pytest.approx(np.sum(models.trapz_weights(x) * y), rel=1e-14)


def test_filter_weight_fold_equals_reference_expression():
    """create_Lnu_grid's W-matrix formulation == the reference's trapz expression (float64)."""
    grid = synthetic.ssp_grid(n_lam=200)
    sed_oracle.ensure_incident(grid)
    F = synthetic.filter_set(grid['lam'], n_filters=3)
    ref = sed_oracle.lnu_grid(grid, F)
    for f in F['filters']:
        w = models.trapz_weights(F[f].lam) * F[f].T / np.trapezoid(F[f].T, x=F[f].lam)
        mine = grid['nebular'] @ w
        np.testing.assert_allclose(mine, ref[f]['nebular'], rtol=1e-13)


def test_observed_geometry_matches_oracle():
    from synthobs_b200.morph import images
    images.filter_info['T.F'] = {'pixel_scale': 0.031}
    cosmo = synthetic.FixedCosmology()
    for kw in (dict(), dict(resampling_factor=2.0), dict(pixel_scale=0.05)):
        o = images.observed('T.F', cosmo, 8.0, 1.7, super_sampling=3, xoffset_pix=0.3, yoffset_pix=-0.2, **kw)
        g = image_oracle.observed_geometry(0.031, cosmo.arcsec_per_kpc, 1.7, super_sampling=3, xoffset_pix=0.3,
                                           yoffset_pix=-0.2, **kw)
        for k in ('pixel_scale', 'pixel_scale_kpc', 'ndim', 'width_arcsec', 'width', 'resolution', 'ndim_super',
                  'xoffset', 'yoffset', 'resampling_factor'):
            assert getattr(o, k) == g[k], k


def test_core_loaders_roundtrip(tmp_path):
    """synthobs.core schema (reference core.py:13-89): write synthetic objects, load them back through the
    reference-named loaders, batch them into CSR."""
    from synthobs_b200 import core, synthetic
    root = str(tmp_path) + '/'
    sizes = {}
    for i, n in enumerate((300, 120, 900)):
        p = synthetic.particles(n, seed=50 + i)
        core.write_object(root + 'snapA/%03d/0' % i, p['X'], p['Y'], p['MetSurfaceDensities'], p['Ages'], p['Metallicities'])
        sizes['snapA/%03d/0' % i] = n
    allobj = core.all_test_data('snapA', path_to_example_data=root)
    assert allobj.N == 3 and list(allobj.M) == sorted(allobj.M)
    first = allobj.next()
    assert first.id == 'snapA/001/0' and len(first.Ages) == 120
    assert np.all(first.Masses == core.PARTICLE_MASS)
    one = core.get_object_data('snapA', '002', '0', data_path=root[:-1])
    ref = synthetic.particles(900, seed=52)
    assert np.array_equal(one.X, ref['X']) and np.array_equal(one.Ages, ref['Ages'])
    td = core.test_data(root + 'snapA/000/0')
    assert len(td.X) == 300
    b = core.csr_batch(allobj.get(i) for i in range(allobj.N))
    assert list(b['offsets']) == [0, 120, 420, 1320]
    np.testing.assert_allclose(b['tauVs_ISM'][120:420], ref_t(300, 50)['tauVs_ISM'], rtol=1e-15)


def ref_t(n, seed):
    from synthobs_b200 import synthetic
    return synthetic.particles(n, seed=seed)


def test_fft_length_rule():
    """transform lengths: >= n, radices 2/3/5/7 only, within 6 % of the smallest such length, a large power of
    two preferred (the measured 1080 -> 1120 case), fixed (no run-time timing)"""
    from synthobs_b200.morph.images import _fft_len, _smooth7
    assert _fft_len(1080) == 1120 and _fft_len(1024) == 1024 and _fft_len(1) == 1 and _fft_len(100) == 100
    for n in list(range(1, 400)) + [513, 1000, 2047, 2049, 4099, 8192 + 56, 12345]:
        m = _fft_len(n)
        first = n
        while not _smooth7(first):
            first += 1
        assert m >= n and _smooth7(m) and m <= first * 1.06 + 1, (n, m, first)
        assert _fft_len(n) == m


def test_ngp_thresholds_refuse_unsorted_axes():
    """np.argmin works for any node order; the threshold form does not and must say so"""
    from synthobs_b200.sed import models
    la = 6.0 + 0.1 * np.arange(51)
    thr = models.ngp_thresholds(la, 6.0)
    assert thr.size == 50 and np.all(np.diff(thr) > 0)
    for bad in (la[::-1], np.r_[la[:10], la[9:]], np.r_[la[:5], np.nan, la[6:]]):
        with pytest.raises(ValueError):
            models.ngp_thresholds(bad, 6.0)


def test_core_sed_known_answers():
    """interrogator's core.sed is not vendored, so its post-processing (examples/SED_spectra.py:103,
    112; SED_Lion.py:38) is restated in synthobs_b200.providers.sed.  The formulas restated are
        get_Lnu(F)[f]  = trapz(lnu T_f, lam) / trapz(T_f, lam)
        get_fnu(cosmo, z): fnu = 1e23 * 1e9 * lnu * (1 + z) / (4 pi dL^2) [nJy] (x IGM transmission), lamz = lam (1 + z)
        get_Fnu(F)[f]  = trapz(fnu T_f, lam_f) / trapz(T_f, lam_f)
        return_log10Q  = log10 integral_{lam < 912} lnu / (h lam) dlam
    and they are pinned here by answers derived by hand: a flat spectrum through any filter, a power law through a
    top hat, the z -> flux scale, and the ionising photon rate of a flat Lyman continuum."""
    from synthobs_b200 import providers
    lam = np.linspace(100.0, 20000.0, 199001)           # 0.1 A steps: trapz error of the power law ~1e-10
    c0 = 3.7e28
    sed = providers.sed(lam)
    F = {'filters': ['th', 'ga']}
    F['th'] = synthetic.Filter(lam, ((lam >= 4000.0) & (lam <= 6000.0)).astype(float))
    F['ga'] = synthetic.Filter(lam, np.exp(-0.5 * ((lam - 9000.0) / 700.0) ** 2))
    sed.lnu = np.full(lam.size, c0)
    L = sed.get_Lnu(F)
    assert L['th'] == pytest.approx(c0, rel=1e-14) and L['ga'] == pytest.approx(c0, rel=1e-14)
    a = -1.5                                            # lnu = lam^a through the top hat [4000, 6000]
    sed.lnu = lam ** a
    exact = (6000.0 ** (a + 1) - 4000.0 ** (a + 1)) / ((a + 1) * 2000.0)
    assert sed.get_Lnu(F)['th'] == pytest.approx(exact, rel=2e-5)      # edge pixels of the top hat: O(h / width)
    z, cosmo = 8.0, synthetic.FixedCosmology()
    sed.lnu = np.full(lam.size, c0)
    fnu = sed.get_fnu(cosmo, z, include_IGM=False)
    scale = 1e23 * 1e9 * (1.0 + z) / (4.0 * np.pi * cosmo.dl_cm ** 2)
    np.testing.assert_allclose(fnu, c0 * scale, rtol=1e-15)
    np.testing.assert_allclose(sed.lamz, lam * 9.0, rtol=0)
    Fz = {'filters': ['th'], 'th': synthetic.Filter(lam * 9.0, F['th'].T)}
    assert sed.get_Fnu(Fz)['th'] == pytest.approx(c0 * scale, rel=1e-14)
    h = 6.626e-27
    top = lam[lam < 912.][-1]                            # strict cut (lam < 912): the integral ends at the last node below
    assert sed.return_log10Q() == pytest.approx(np.log10(c0 / h * np.log(top / 100.0)), abs=1e-7)   # trapz of 1/lam on 0.1 A steps
    # IGM: transparent redward of Lyman alpha in the observed frame, opaque far blueward
    t = providers.madau(np.array([500.0 * 9, 1300.0 * 9, 20000.0]), 8.0)
    assert t[1] == 1.0 and t[2] == 1.0 and t[0] < 1e-3


def test_image_reduce_mode_env(monkeypatch):
    """dist.reduce_mode: SO_PEER_REDUCE = 1 / 0 force the peer-memory exchange / ncclAllReduce, anything else measures"""
    from synthobs_b200 import dist as sdist
    monkeypatch.delenv('SO_PEER_REDUCE', raising=False)
    assert sdist.reduce_mode() == 'auto' and sdist.peer_reduce_enabled()
    monkeypatch.setenv('SO_PEER_REDUCE', '1')
    assert sdist.reduce_mode() == 'peer'
    monkeypatch.setenv('SO_PEER_REDUCE', '0')
    assert sdist.reduce_mode() == 'nccl' and not sdist.peer_reduce_enabled()
    # every timed mode is preceded by an untimed call of the same mode
    assert sdist.REDUCE_TRIALS[0::2] == sdist.REDUCE_TRIALS[1::2] and set(sdist.REDUCE_TRIALS) == {'peer', 'nccl'}
    assert sdist.SHARE_ADAPT_CALLS >= 2
