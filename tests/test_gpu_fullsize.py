"""GPU: BASELINE.json's full sizes through size-independent properties (the oracle would take
minutes to hours there): linearity, permutation invariance, batch == concatenation, flux
conservation, count conservation, determinism, rebin sum preservation."""

import contextlib
import io

import numpy as np
import pytest

from synthobs_b200 import synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def model_full():
    from synthobs_b200.sed import models
    grid = synthetic.ssp_grid(n_lam=9000, seed=0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = models.define_model(grid)
    m.dust_ISM = ('simple', {'slope': -1.0})
    m.dust_BC = ('simple', {'slope': -1.0})
    F = synthetic.filter_set(m.lam * 9.0, n_filters=10, kind='gauss', lo=12000., hi=50000., prefix='J.')
    m.create_Fnu_grid(F, 8.0, synthetic.FixedCosmology())
    return m, F


def test_config2_sed_properties(model_full):
    """config[1]: 1e5 particles, 51x13x9000 grid, ISM+BC dust, 10 filters"""
    from synthobs_b200.sed import models
    model, F = model_full
    p = synthetic.particles(100000, seed=41, lognormal_mass=True)
    args = (p['Masses'], p['Ages'], p['Metallicities'])
    kw = dict(tauVs_ISM=p['tauVs_ISM'], tauVs_BC=p['tauVs_BC'], fesc=0.1)
    f = models.generate_Fnu(model, F, *args, **kw)
    base = np.array([f[k] for k in F['filters']])
    assert np.all(np.isfinite(base)) and np.all(base > 0)
    # linearity in mass
    f2 = models.generate_Fnu(model, F, 2.0 * p['Masses'], p['Ages'], p['Metallicities'], **kw)
    np.testing.assert_allclose([f2[k] for k in F['filters']], 2 * base, rtol=2e-6)
    # permutation invariance
    perm = np.random.default_rng(0).permutation(100000)
    fp = models.generate_Fnu(model, F, p['Masses'][perm], p['Ages'][perm], p['Metallicities'][perm],
                             tauVs_ISM=p['tauVs_ISM'][perm], tauVs_BC=p['tauVs_BC'][perm], fesc=0.1)
    np.testing.assert_allclose([fp[k] for k in F['filters']], base, rtol=2e-6)
    # split into 7 "galaxies": batch rows add up to the whole
    off = np.array([0, 1, 1, 5000, 33333, 33334, 99999, 100000])
    rows = models.generate_Fnu_batch(model, F, *args, off, **kw)
    np.testing.assert_allclose(rows.sum(axis=0), base, rtol=2e-6)
    assert np.all(rows[1] == 0)          # empty galaxy
    # sum of the per-particle arrays == totals (models.py:83)
    arrs = models.generate_Fnu_arrays(model, F, *args, **kw)
    np.testing.assert_allclose([arrs[k].sum() for k in F['filters']], base, rtol=2e-6)
    # more dust -> less flux, no dust -> upper bound
    fd = models.generate_Fnu(model, F, *args, tauVs_ISM=3 * p['tauVs_ISM'], tauVs_BC=p['tauVs_BC'], fesc=0.1)
    f0 = models.generate_Fnu(model, F, *args, fesc=0.1)
    assert all(fd[k] < f[k] < f0[k] for k in F['filters'])


def test_config2_spectra_properties(model_full):
    from synthobs_b200.sed import models
    model, F = model_full
    p = synthetic.particles(100000, seed=42)
    args = (p['Masses'], p['Ages'], p['Metallicities'])
    with contextlib.redirect_stdout(io.StringIO()):
        o = models.generate_SED(model, *args, tauVs_ISM=p['tauVs_ISM'], tauVs_BC=p['tauVs_BC'], fesc=0.0)
        nod = models.generate_SED(model, *args, fesc=0.0)      # tau arrays absent -> exp(-0) = 1
    assert o.total.lnu.shape == (9000,)
    # attenuation ordering: total <= BC <= intrinsic ; no_HII_no_BC <= stellar
    tol = 1e-5 * o.intrinsic.lnu.max()
    assert np.all(o.total.lnu <= o.BC.lnu + tol) and np.all(o.BC.lnu <= o.intrinsic.lnu + tol)
    assert np.all(o.no_HII_no_BC.lnu <= o.stellar.lnu + tol)
    # dust-free limits reduce to the tensor-core contractions
    np.testing.assert_allclose(nod.total.lnu, nod.intrinsic.lnu, rtol=2e-5, atol=tol)
    np.testing.assert_allclose(nod.no_HII_no_BC.lnu, nod.stellar.lnu, rtol=2e-5, atol=tol)
    np.testing.assert_allclose(o.stellar.lnu, nod.stellar.lnu, rtol=1e-6)
    # the Lnu-grid path and the spectrum path agree in the dust-free case (examples/SED_luminosities.py:18)
    Fr = synthetic.filter_set(model.lam, n_filters=6)
    model.create_Lnu_grid(Fr)
    keep_I, keep_B = model.dust_ISM, model.dust_BC
    model.dust_ISM = model.dust_BC = False
    try:
        L = models.generate_Lnu(model, Fr, *args)
    finally:
        model.dust_ISM, model.dust_BC = keep_I, keep_B
    for f in Fr['filters']:
        band = np.trapezoid(nod.intrinsic.lnu * Fr[f].T, x=Fr[f].lam) / np.trapezoid(Fr[f].T, x=Fr[f].lam)
        assert band == pytest.approx(L[f], rel=1e-4)


def test_config3_imaging_properties():
    """config[2]: 1e6 particles, 512x512 observed frame, super-sampling 2, adaptive k=8, PSF"""
    import torch
    from synthobs_b200.morph import images
    p = synthetic.particles(1000000, seed=43, scale_kpc=3.0)
    filters = ['P.A', 'P.B']
    for f in filters:
        images.filter_info[f] = {'pixel_scale': 0.031}
    cosmo = synthetic.FixedCosmology()
    obs = [images.observed(f, cosmo, 8.0, 512 * 0.031 + 1e-9, smoothing=('adaptive', 8.), PSF=synthetic.GaussianPSF(2.0),
                           super_sampling=2) for f in filters]
    assert obs[0].ndim == 512 and obs[0].ndim_super == 1024
    X = torch.from_numpy(p['X'] - np.median(p['X'])).cuda()
    Y = torch.from_numpy(p['Y'] - np.median(p['Y'])).cuda()
    w = np.stack([p['Masses'], p['Masses'] * np.random.default_rng(1).random(1000000)])
    L = torch.from_numpy(w).cuda()
    r1 = images.particle_device_multi(obs, X, Y, L)
    r2 = images.particle_device_multi(obs, X, Y, L)
    assert torch.equal(r1['data'], r2['data']) and torch.equal(r1['super_no_PSF'], r2['super_no_PSF'])   # deterministic
    sel = r1['stats']['selected']
    assert sel == 1000000 and r1['stats']['dropped'] == 0
    for c in range(2):
        tot = w[c].sum()
        assert float(r1['super_no_PSF'][c].double().sum()) == pytest.approx(tot, rel=5e-6)     # flux conservation
        assert float(r1['no_PSF'][c].double().sum()) == pytest.approx(tot, rel=5e-6)           # rebin keeps sums
        assert float(r1['data'][c].double().sum()) == pytest.approx(tot, rel=2e-3)             # PSF wings leave the frame
        assert float(r1['super_no_PSF'][c].min()) >= 0.0
    # NGP products: counts conserve the particle number
    h = images.deposit_device(X, Y, L[:1], obs[0].resolution, 1024, adaptive_k=None)
    assert float(h['hist'].double().sum()) == 1000000
    assert float(h['simple'][0].double().sum()) == pytest.approx(w[0].sum(), rel=1e-6)
    # linearity in the weights
    r3 = images.particle_device_multi(obs, X, Y, 3.0 * L)
    assert torch.allclose(r3['no_PSF'], 3.0 * r1['no_PSF'], rtol=1e-5, atol=1e-7 * float(r1['no_PSF'].max()))


def test_large_image_flux_and_odd_sizes():
    """a 4099x4099 image (not a multiple of the 16-pixel sub-tile) from 2e6 particles"""
    import torch
    from synthobs_b200 import device as dv
    from synthobs_b200.morph import images
    p = synthetic.particles(2000000, seed=44, scale_kpc=6.0)
    r = images.deposit_device(dv.to_dev(p['X']), dv.to_dev(p['Y']), dv.to_dev(p['Masses']), 0.01, 4099, adaptive_k=5,
                              want_ngp=False)
    width = 4099 * 0.01
    sel = (np.abs(p['X']) < width / 2) & (np.abs(p['Y']) < width / 2)
    assert r['stats']['selected'] == sel.sum()
    assert float(r['data'].double().sum()) == pytest.approx(p['Masses'][sel].sum(), rel=5e-6)


def test_config4_batch_properties(model_full):
    """config[3] (one GPU's share, scaled to 300 galaxies of 1e3..1e5 particles): batched SEDs, emission lines and
    images.  Batch rows == the single-call result on the concatenation; flux and counts conserved per object; the
    batch image of a sampled object == core() on that object alone."""
    from synthobs_b200.sed import models
    from synthobs_b200.morph import images
    model, F = model_full
    b = synthetic.galaxy_batch(300, n_min=1000, n_max=100000, seed=77)
    off = b['offsets']
    n = b['Masses'].size
    assert n > 4e6
    args = (b['Masses'], b['Ages'], b['Metallicities'])
    kw = dict(tauVs_ISM=b['tauVs_ISM'], tauVs_BC=b['tauVs_BC'])
    rows = models.generate_Fnu_batch(model, F, *args, off, **kw)
    tot = models.generate_Fnu(model, F, *args, **kw)
    np.testing.assert_allclose(rows.sum(axis=0), [tot[k] for k in F['filters']], rtol=2e-6)
    g = 123
    s, e = off[g], off[g + 1]
    one = models.generate_Fnu(model, F, b['Masses'][s:e], b['Ages'][s:e], b['Metallicities'][s:e],
                              tauVs_ISM=b['tauVs_ISM'][s:e], tauVs_BC=b['tauVs_BC'][s:e])
    np.testing.assert_allclose(rows[g], [one[k] for k in F['filters']], rtol=2e-6)
    # emission lines, batched: rows add up to the whole catalogue
    with contextlib.redirect_stdout(io.StringIO()):
        el = models.EmissionLines(synthetic.lines_grid(seed=0), dust_ISM=('simple', {'slope': -1.0}),
                                  dust_BC=('simple', {'slope': -1.0}), verbose=False)
        lb = el.get_line_luminosities_batch(['HI6563', 'OIII4959,OIII5007'], *args, off, **kw)
        lt = el.get_line_luminosities(['HI6563', 'OIII4959,OIII5007'], *args, **kw)
    for line in lb:
        for q in ('luminosity', 'continuum'):
            np.testing.assert_allclose(lb[line][q].sum(), lt[line][q], rtol=2e-6)
    # images: 300 objects, 100 x 100, adaptive smoothing, one pass
    L = b['Masses'] * (0.5 + np.random.default_rng(3).random(n))
    img = images.core_batch(b['X'], b['Y'], L, off, resolution=0.1, ndim=100, smoothing=('adaptive', 16))
    assert img.data.shape == (300, 100, 100)
    sel = (np.abs(b['X']) < 5.0) & (np.abs(b['Y']) < 5.0)
    flux = np.add.reduceat(np.where(sel, L, 0.0), off[:-1])
    cnt = np.add.reduceat(sel.astype(np.float64), off[:-1])
    np.testing.assert_allclose(img.data.reshape(300, -1).sum(axis=1), flux, rtol=1e-5)
    assert np.array_equal(img.hist.reshape(300, -1).sum(axis=1), cnt)
    one = images.core(b['X'][s:e], b['Y'][s:e], L[s:e], resolution=0.1, ndim=100, smoothing=('adaptive', 16))
    assert np.array_equal(img.hist[g], one.hist)
    assert np.abs(img.data[g] - one.data).max() <= 2e-6 * one.data.max()


def test_config5_huge_image_properties():
    """config[4]: 1e8 particles onto 8192 x 8192, adaptive smoothing (single rank of the particle-share path):
    flux and selected-particle count conserved, non-negative, and the two halves of the Morton-ordered particle
    shares (what two ranks would deposit) add up to the whole image."""
    import torch
    from synthobs_b200 import device as dv
    from synthobs_b200.morph import images
    dev = dv.device()
    N, ndim, S = 100000000, 8192, 10.0
    g = torch.Generator(device=dev)
    g.manual_seed(99)
    X = S * torch.randn(N, dtype=torch.float64, device=dev, generator=g)
    Y = S * torch.randn(N, dtype=torch.float64, device=dev, generator=g)
    m = N // 20
    X[:m] = 3.0 + 0.5 * torch.randn(m, dtype=torch.float64, device=dev, generator=g)      # a dense clump
    Y[:m] = -7.0 + 0.5 * torch.randn(m, dtype=torch.float64, device=dev, generator=g)
    L = 8.47e5 * (0.5 + torch.rand(N, dtype=torch.float64, device=dev, generator=g))
    res = 8.0 * S / ndim
    width, G = images.pixel_grid(res, ndim)
    ix, iy = images.ngp_index_device(X, Y, dv.to_dev(G), ndim, width / 2.)
    sel = ix >= 0
    flux = float(L[sel].sum())
    parts = []
    for part in range(2):
        dk, pids = images.knn_device_part(X, Y, ix, 8, width / 2., part, 2)
        r = images.deposit_device(X[pids], Y[pids], L[pids].reshape(1, -1), res, ndim, adaptive_k=8, want_ngp=False,
                                  dk=dk[pids])
        parts.append((r['data'][0], int(pids.numel()), r['stats']))
        del dk, pids, r
    assert parts[0][1] + parts[1][1] == int(sel.sum())
    img = parts[0][0] + parts[1][0]
    assert float(img.min()) >= 0.0
    assert abs(float(img.to(torch.float64).sum()) / flux - 1.0) < 2e-6
    assert parts[0][2]['dropped'] == 0 and parts[1][2]['dropped'] == 0
    # the shares are spatially compact: each touches clearly fewer pixels than the whole image
    nz0, nz1, nz = (int((parts[0][0] > 0).sum()), int((parts[1][0] > 0).sum()), int((img > 0).sum()))
    assert nz0 < 0.8 * nz and nz1 < 0.8 * nz
