"""GPU parity: the CUDA imaging path (through the C-ABI) against the golden vectors written
by the unmodified reference and against the oracle.

Tolerances: NGP pixel indices, counts and particle->tile assignment bit-exact; images within
1e-5 of the image peak per pixel and 1e-5 relative in L2 (FP32; see DESIGN.md for why a
truncated-support gather cannot be per-pixel relative in empty regions); total flux 1e-6."""

import numpy as np
import pytest

from oracle import image_oracle
from synthobs_b200 import synthetic

import _cases

pytestmark = pytest.mark.gpu


def assert_image_close(a, b, tol=1e-5):
    peak = np.abs(b).max()
    assert a.shape == b.shape
    assert np.abs(a - b).max() <= tol * peak, (np.abs(a - b).max() / peak)
    assert np.linalg.norm(a - b) <= tol * np.linalg.norm(b)


@pytest.mark.parametrize('tag', _cases.IMG_CASES)
def test_core_matches_reference(tag):
    from synthobs_b200.morph import images
    g = _cases.load('img_' + tag)
    sm = _cases.smoothing_of(g)
    img = images.core(g['X'].copy(), g['Y'].copy(), g['L'].copy(), resolution=float(g['resolution']),
                      ndim=int(g['ndim']), smoothing=sm)
    assert img.hist.dtype == np.float64
    assert np.array_equal(img.hist, g['hist'])
    np.testing.assert_allclose(img.simple, g['simple'], rtol=2e-7, atol=0)
    assert_image_close(img.data, g['data'])
    assert img.data.sum() == pytest.approx(g['data'].sum(), rel=2e-6)
    if not sm:
        assert img.data is img.simple          # images.py:103


@pytest.mark.parametrize('tag', ['adaptive16', 'adaptive_odd', 'adaptive_kbig', 'adaptive_underflow'])
def test_full_support_is_per_pixel_relative(tag):
    """SUPPORT_SIGMAS = 0 evaluates every pixel like the reference: per-pixel relative 1e-5
    wherever the value is representable in FP32 relative to the peak."""
    from synthobs_b200.morph import images
    g = _cases.load('img_' + tag)
    old = images.SUPPORT_SIGMAS
    images.SUPPORT_SIGMAS = 0.0
    try:
        img = images.core(g['X'].copy(), g['Y'].copy(), g['L'].copy(), resolution=float(g['resolution']),
                          ndim=int(g['ndim']), smoothing=_cases.smoothing_of(g))
    finally:
        images.SUPPORT_SIGMAS = old
    ref = g['data']
    big = ref > 1e-25 * ref.max()
    np.testing.assert_allclose(img.data[big], ref[big], rtol=2e-5)
    assert_image_close(img.data, ref)


def test_pixel_indices_bit_exact():
    import torch
    from synthobs_b200.morph import images
    from synthobs_b200 import device as dv
    g = _cases.load('img_indices')
    width, G = images.pixel_grid(float(g['resolution']), int(g['ndim']))
    x = g['x']
    ix, iy = images.ngp_index_device(dv.to_dev(x), dv.to_dev(np.full_like(x, 0.123)), dv.to_dev(G), int(g['ndim']),
                                     width / 2.)
    assert np.array_equal(ix.cpu().numpy(), g['idx'])
    # large random + midpoints on a big grid vs the oracle's literal argmin
    width, G = images.pixel_grid(0.013, 1024)
    rng = np.random.default_rng(0)
    mids = 0.5 * (G[:-1] + G[1:])
    xs = np.concatenate([rng.uniform(-width / 2, width / 2, 200000), mids, np.nextafter(mids, 9), np.nextafter(mids, -9),
                         [width / 2, -width / 2, np.nextafter(width / 2, 0), np.nan]])
    ys = np.zeros_like(xs)
    ix, iy = images.ngp_index_device(dv.to_dev(xs), dv.to_dev(ys), dv.to_dev(G), 1024, width / 2.)
    sel = image_oracle.select(xs, ys, width)
    ref = np.full(xs.size, -1)
    ref[sel] = image_oracle.ngp_pixels(G, xs[sel])
    assert np.array_equal(ix.cpu().numpy(), ref)


@pytest.mark.parametrize('n,k', [(3000, 16), (20000, 8), (500, 1), (64, 64), (40, 50)])
def test_knn_distance_matches_ckdtree(n, k):
    from synthobs_b200.morph import images
    from synthobs_b200 import device as dv
    p = synthetic.particles(n, seed=n + k)
    X, Y = p['X'], p['Y']
    X[:5] = X[5:10]          # exact duplicates
    Y[:5] = Y[5:10]
    hw = 4.0
    width, G = images.pixel_grid(2 * hw / 100, 100)
    Xd, Yd = dv.to_dev(X), dv.to_dev(Y)
    ix, _ = images.ngp_index_device(Xd, Yd, dv.to_dev(G), 100, hw)
    dk = images.knn_device(Xd, Yd, ix, k, hw).cpu().numpy()
    sel = image_oracle.select(X, Y, 2 * hw)
    ref = image_oracle.kth_neighbour_distance(X[sel], Y[sel], k) if k <= sel.sum() else np.full(sel.sum(), np.inf)
    assert np.array_equal(dk[~sel], np.zeros((~sel).sum()))
    np.testing.assert_allclose(dk[sel], ref, rtol=4e-16, atol=0)


def test_tile_assignment_bit_exact():
    from synthobs_b200.morph import images
    from synthobs_b200 import device as dv
    g = _cases.load('img_adaptive8f')
    ndim, res = 200, 0.02
    X, Y, L = g['X'], g['Y'], g['L']
    r = images.deposit_device(dv.to_dev(X), dv.to_dev(Y), dv.to_dev(L), res, ndim, adaptive_k=8, return_pairs=True,
                              exact_dk=True)
    o = image_oracle.core(X, Y, L, resolution=res, ndim=ndim, smoothing=('adaptive', 8), return_aux=True)
    sel = o['aux']['sel']
    ix = r['ix'].cpu().numpy()
    iy = r['iy'].cpu().numpy()
    assert np.array_equal(ix >= 0, sel)
    assert np.array_equal(ix[sel], o['aux']['i']) and np.array_equal(iy[sel], o['aux']['j'])
    G = o['aux']['G']
    pitch = (G[-1] - G[0]) / (ndim - 1)
    dk = r['dk'].cpu().numpy()[sel]
    np.testing.assert_allclose(dk, o['aux']['dk'], rtol=4e-16)
    W = image_oracle.support_halfwidth(dk, pitch, images.SUPPORT_SIGMAS)
    tiles, pids = image_oracle.tile_pairs(o["aux"]["i"], o["aux"]["j"], W, ndim, 16)
    gid = np.nonzero(sel)[0]
    assert np.array_equal(r['tiles'].cpu().numpy(), tiles)
    assert np.array_equal(r['pids'].cpu().numpy(), gid[pids])
    assert r['stats']['pairs'] == tiles.size
    assert_image_close(r['data'][0].cpu().numpy().astype(np.float64), o['data'])


def test_oracle_parity_larger_seeded_case_and_determinism():
    """1e4 particles, 100x100 (BASELINE config 1 shape) vs the oracle; twice -> bitwise equal"""
    from synthobs_b200.morph import images
    p = synthetic.particles(10000, seed=1)
    L = p['Masses']
    a = images.core(p['X'], p['Y'], L, resolution=0.1, ndim=100, smoothing=('adaptive', 16))
    b = images.core(p['X'], p['Y'], L, resolution=0.1, ndim=100, smoothing=('adaptive', 16))
    assert np.array_equal(a.data, b.data)
    o = image_oracle.core(p['X'], p['Y'], L, resolution=0.1, ndim=100, smoothing=('adaptive', 16), return_aux=True)
    assert np.array_equal(a.hist, o['hist'])
    assert_image_close(a.data, o['data'])
    assert a.data.sum() == pytest.approx(L[o['aux']['sel']].sum(), rel=2e-6)      # flux conservation


def test_multi_channel_and_empty():
    import torch
    from synthobs_b200.morph import images
    from synthobs_b200 import device as dv
    p = synthetic.particles(4000, seed=5)
    L = np.stack([p['Masses'], p['Masses'] * np.random.default_rng(1).random(4000), np.zeros(4000)])
    r = images.deposit_device(dv.to_dev(p['X']), dv.to_dev(p['Y']), dv.to_dev(L), 0.05, 160, adaptive_k=8)
    for c in range(3):
        o = image_oracle.core(p['X'], p['Y'], L[c], resolution=0.05, ndim=160, smoothing=('adaptive', 8))
        if c < 2:
            assert_image_close(r['data'][c].cpu().numpy().astype(np.float64), o['data'])
        else:
            assert float(r['data'][c].abs().max()) == 0.0
    # channel groups of 8 / 4 / 2 / 1 with and without 16-byte aligned weight rows (12 = 8 + 4, 13 = 8 + 4 + 1, 7 = 4 + 2 + 1),
    # odd image size: every channel is a multiple of the first one (the deposit is linear in the weights)
    o = image_oracle.core(p['X'], p['Y'], p['Masses'], resolution=0.05, ndim=157, smoothing=('adaptive', 8))
    for C in (12, 13, 7):
        f = 0.25 + np.arange(C)
        r = images.deposit_device(dv.to_dev(p['X']), dv.to_dev(p['Y']), dv.to_dev(p['Masses'][None, :] * f[:, None]), 0.05, 157,
                                  adaptive_k=8)
        d = r['data'].cpu().numpy().astype(np.float64)
        assert_image_close(d[0] / f[0], o['data'])
        for c in range(1, C):
            assert np.abs(d[c] / f[c] - d[0] / f[0]).max() <= 2e-6 * np.abs(o['data']).max(), (C, c)
    # nothing selected
    img = images.core(np.array([100.0, -100.0]), np.array([0.0, 0.0]), np.array([1.0, 1.0]), ndim=30,
                      smoothing=('adaptive', 4))
    assert img.data.shape == (30, 30) and img.data.sum() == 0 and img.hist.sum() == 0
    img = images.core(np.zeros(0), np.zeros(0), np.zeros(0), ndim=30, smoothing=('adaptive', 4))
    assert img.data.sum() == 0


def test_rebin_and_convolve():
    from synthobs_b200.morph import images
    rng = np.random.default_rng(2)
    a = rng.random((120, 120))
    np.testing.assert_allclose(images.rebin(a, (40, 40)), image_oracle.rebin(a, (40, 40)), rtol=1e-6)
    assert images.rebin(a, (40, 40)).sum() == pytest.approx(a.sum(), rel=1e-6)
    for nk in (120, 121, 33):
        k = rng.random((nk, nk))
        ref = image_oracle.convolve_fft_same(a, k)
        got = images.convolve_fft(a, k)
        assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()


def test_observed_matches_reference():
    from synthobs_b200.morph import images
    g = _cases.load('img_observed')
    cosmo = synthetic.FixedCosmology(arcsec_per_kpc=float(g['arcsec_per_kpc']))
    filters = ['JWST.NIRCAM.F150W', 'JWST.NIRCAM.F356W']
    images.filter_info.update({'JWST.NIRCAM.F150W': {'pixel_scale': 0.031}, 'JWST.NIRCAM.F356W': {'pixel_scale': 0.063}})
    PSFs = {f: synthetic.GaussianPSF(fw) for f, fw in zip(filters, g['fwhm'])}
    X, Y = g['X'].copy(), g['Y'].copy()
    o = images.observed(filters[0], cosmo, float(g['z']), float(g['width_arcsec']), smoothing=('adaptive', 8.),
                        PSF=PSFs[filters[0]], super_sampling=int(g['super_sampling']), xoffset_pix=0.25,
                        yoffset_pix=-0.4).particle(X, Y, g['L_0'])
    assert np.array_equal(X, g['obs_X_after']) and np.array_equal(Y, g['obs_Y_after'])   # in-place recentring
    assert o.img.pixel_scale == float(g['obs_pixel_scale'])
    assert_image_close(o.super.no_PSF, g['obs_super_no_PSF'])
    assert_image_close(o.super.data, g['obs_super_data'])
    assert_image_close(o.img.no_PSF, g['obs_no_PSF'])
    assert_image_close(o.img.data, g['obs_data'])
    X, Y = g['X'].copy(), g['Y'].copy()
    L = {f: g['L_%d' % i] for i, f in enumerate(filters)}
    IMGs = images.particle(X, Y, L, filters, cosmo, float(g['z']), float(g['width_arcsec']),
                           smoothing=('adaptive', 8.), PSFs=PSFs, super_sampling=int(g['super_sampling']))
    assert np.array_equal(X, g['multi_X_after'])
    for i, f in enumerate(filters):
        assert_image_close(IMGs[f].data, g['multi_%d_data' % i])
        assert_image_close(IMGs[f].no_PSF, g['multi_%d_no_PSF' % i])


def test_particle_fast_path_equals_per_filter_loop():
    """images.particle images the filters of one geometry together (one k-NN, one binning, multi-channel deposit,
    batched FFT).  Against the reference's own structure -- a loop of observed(...).particle per filter
    (images.py:277-284) -- the caller's X, Y must end up bit-identical (one in-place recentring per filter) and the
    images equal; random sub-pixel offsets (global np.random, images.py:265-267), two pixel scales, even N (the
    median residue of the second recentring is then not zero)."""
    import torch
    from synthobs_b200.morph import images
    p = synthetic.particles(6000, seed=31, scale_kpc=1.0)
    filters = ['T.A', 'T.B', 'T.C', 'T.D']
    scales = {'T.A': 0.031, 'T.B': 0.063, 'T.C': 0.031, 'T.D': 0.031}
    for f in filters:
        images.filter_info[f] = {'pixel_scale': scales[f]}
    cosmo = synthetic.FixedCosmology()
    PSFs = {f: synthetic.GaussianPSF(1.4 + 0.3 * i) for i, f in enumerate(filters)}
    rng = np.random.default_rng(2)
    L = {f: p['Masses'] * (0.5 + rng.random(6000)) for f in filters}
    kw = dict(smoothing=('adaptive', 8.), super_sampling=2)
    X1, Y1 = p['X'].copy(), p['Y'].copy()
    np.random.seed(11)
    got = images.particle(X1, Y1, L, filters, cosmo, 8.0, 3.0, PSFs=PSFs, offsets=True, **kw)
    # the reference's loop, spelled out with observed.particle
    X2, Y2 = p['X'].copy(), p['Y'].copy()
    np.random.seed(11)
    bx = np.random.random() - 0.5
    np.random.random()
    mps = max(scales.values())
    for f in filters:
        off = bx * (mps / scales[f])
        ref = images.observed(f, cosmo, 8.0, 3.0, PSF=PSFs[f], xoffset_pix=off, yoffset_pix=off, **kw).particle(X2, Y2, L[f])
        assert got[f].pixel_scale == ref.img.pixel_scale and got[f].data.shape == ref.img.data.shape
        assert_image_close(got[f].data, ref.img.data, 2e-6)
        assert_image_close(got[f].no_PSF, ref.img.no_PSF, 2e-6)
    assert np.array_equal(X1, X2) and np.array_equal(Y1, Y2)
    # CUDA tensors in -> CUDA tensors out, same values, tensors recentred in place
    Xd, Yd = torch.from_numpy(p['X']).cuda(), torch.from_numpy(p['Y']).cuda()
    np.random.seed(11)
    gd = images.particle(Xd, Yd, {f: torch.from_numpy(L[f]).cuda() for f in filters}, filters, cosmo, 8.0, 3.0,
                         PSFs=PSFs, offsets=True, **kw)
    assert gd['T.B'].data.is_cuda and np.array_equal(Xd.cpu().numpy(), X2)
    for f in filters:
        assert np.array_equal(gd[f].data.double().cpu().numpy(), got[f].data)


@pytest.mark.parametrize('n_parts', [2, 5, 8])
def test_share_plan_keeps_knn_exact(n_parts):
    """so_share_plan: the shares are disjoint, cover every selected particle, are balanced, and the k-th neighbour
    distance of every OWN particle computed among the LOCAL particles only (own + halo) is bit-identical to the
    one computed among all particles -- clustered catalogue, sparse outskirts, k larger than many coarse cells hold"""
    import torch
    from synthobs_b200 import device as dv
    from synthobs_b200.morph import images
    p = synthetic.particles(400000, seed=77, scale_kpc=2.0)
    X, Y = dv.to_dev(p['X']), dv.to_dev(p['Y'])
    hw, k = 9.0, 12
    width, G, Gd = images.pixel_grid_device(2 * hw / 2000, 2000)
    ix, _ = images.ngp_index_device(X, Y, Gd, 2000, hw)
    full = images.knn_device(X, Y, ix, k, hw)
    sel = ix >= 0
    owned = torch.zeros_like(sel, dtype=torch.int32)
    for part in range(n_parts):
        for cb in (None, 7):
            flags, counts = images.share_plan_device(X, Y, hw, k, part, n_parts, cb=cb)
            loc = torch.nonzero(flags).reshape(-1)
            own = (flags & 2) != 0
            assert int(counts[0]) == loc.numel() and int(counts[1]) == int(own.sum())
            assert bool(((flags & 1) != 0)[own].all()) and not bool(own[~sel].any())
            ixl, _ = images.ngp_index_device(X[loc], Y[loc], Gd, 2000, hw)
            dkl = images.knn_device_masked(X[loc], Y[loc], ixl, k, hw, flags[loc])
            o = own[loc]
            assert torch.equal(dkl[o], full[loc][o])
            if cb is None:
                owned += own.to(torch.int32)
                # equal WORK up to one coarse cell (sparse cells weigh three times): the counts stay within a factor
                assert 0.3 < int(own.sum()) / (int(sel.sum()) / n_parts) < 2.0
    assert torch.equal(owned, sel.to(torch.int32))


def test_image_huge_single_rank_equals_core():
    """particle-block path (BASELINE config 5) on one rank == images.core adaptive"""
    from synthobs_b200 import dist as sdist
    from synthobs_b200.morph import images
    p = synthetic.particles(30000, seed=12)
    img = sdist.image_huge(p['X'], p['Y'], p['Masses'], 0.02, 400, 8).cpu().numpy().astype(np.float64)
    o = image_oracle.core(p['X'], p['Y'], p['Masses'], resolution=0.02, ndim=400, smoothing=('adaptive', 8))
    assert_image_close(img, o['data'])


def test_analytic_sources_match_reference():
    """observed.Sersic / Sersic() / point() vs the golden vectors of the
    unmodified reference: Sersic profile kernel (float64 evaluation), cuFFT PSF step, rebin."""
    from synthobs_b200.morph import images
    images.filter_info.update({'JWST.NIRCAM.F150W': {'pixel_scale': 0.031}, 'JWST.NIRCAM.F356W': {'pixel_scale': 0.063}})
    g = _cases.load('img_analytic')
    cosmo = synthetic.FixedCosmology()
    p = dict(zip(('r_eff', 'n', 'ellip', 'theta'), [float(v) for v in g['sersic_p']]))
    o = images.observed('JWST.NIRCAM.F150W', cosmo, 8.0, 0.8, resampling_factor=2, PSF=synthetic.GaussianPSF(2.1),
                        super_sampling=3, xoffset_pix=0.3, yoffset_pix=-0.2).Sersic(3.5e28, p)
    assert_image_close(o.super.no_PSF, g['ser_super_no_PSF'])
    assert_image_close(o.super.data, g['ser_super_data'])
    assert_image_close(o.img.no_PSF, g['ser_no_PSF'])
    assert_image_close(o.img.data, g['ser_data'])
    np.testing.assert_allclose(o.super.no_PSF, g['ser_super_no_PSF'], rtol=2e-7)      # float64 profile, one rounding
    assert o.img.pixel_scale == float(g['ser_pixel_scale']) and o.img.pixel_scale_kpc == float(g['ser_pixel_scale_kpc'])
    filters = ['JWST.NIRCAM.F150W', 'JWST.NIRCAM.F356W']
    PSFs = {'JWST.NIRCAM.F150W': synthetic.GaussianPSF(2.1), 'JWST.NIRCAM.F356W': synthetic.GaussianPSF(1.8)}
    L = {'JWST.NIRCAM.F150W': 1.0e28, 'JWST.NIRCAM.F356W': 2.5e28}
    IMGs = images.Sersic(L, {'r_eff': 0.5, 'n': 4.0, 'ellip': 0.0, 'theta': 0.0}, filters, cosmo, 8.0, 1.2, PSFs=PSFs,
                         super_sampling=2)
    for fi, f in enumerate(filters):
        assert_image_close(IMGs[f].data, g['sermulti_%d_data' % fi])
        assert_image_close(IMGs[f].no_PSF, g['sermulti_%d_no_PSF' % fi])
    psf = synthetic.airy_like_psf(ndim=101, sub=5, fwhm_pix=2.0, pixel_scale=0.031)
    o = images.point(7.0e3, 'JWST.NIRCAM.F150W', 0.55, PSF=psf, super_sampling=5, xoffset_pix=0.21, yoffset_pix=-0.37)
    assert np.array_equal(o.super.hist, g['pt_super_hist'])                            # the source's pixel: bit-exact
    assert np.array_equal(o.super.simple != 0, g['pt_super_simple'] != 0)
    assert_image_close(o.super.data, g['pt_super_data'])
    assert_image_close(o.img.no_PSF, g['pt_no_PSF'])
    assert_image_close(o.img.data, g['pt_data'])
    assert o.img.data.sum() == pytest.approx(7.0e3, rel=1e-6)
    assert o.img.ndim == int(g['pt_ndim']) and o.img.pixel_scale == float(g['pt_pixel_scale'])
    psf2 = synthetic.airy_like_psf(ndim=75, sub=5, fwhm_pix=1.4, pixel_scale=0.063, seed=4)
    o = images.point(2.0, 'JWST.NIRCAM.F356W', 1.5, resampling_factor=2, super_sampling=3, PSF=psf2)
    assert_image_close(o.super.data, g['pt2_super_data'])
    assert_image_close(o.img.data, g['pt2_data'])
    # points(): common random offset from the global np.random, as upstream (images.py:421-422)
    np.random.seed(3)
    xo, yo = np.random.random() - 0.5, np.random.random() - 0.5
    np.random.seed(3)
    P = images.points({'JWST.NIRCAM.F150W': 5.0, 'JWST.NIRCAM.F356W': 9.0}, filters, 0.9,
                      PSFs={'JWST.NIRCAM.F150W': psf, 'JWST.NIRCAM.F356W': psf2})
    r = image_oracle.point_source(9.0, 0.063, 0.9, psf2, xoffset_pix=xo, yoffset_pix=yo)
    assert_image_close(P['JWST.NIRCAM.F356W'].data, r['data'])


def test_median_matches_numpy_bitwise():
    """so_median_f64 (radix selection) == np.median bit for bit: odd / even n, duplicates around the
    middle, negative values, batched rows"""
    import torch
    from synthobs_b200.morph import images
    rng = np.random.default_rng(17)
    cases = [rng.normal(0, 3, 1001), rng.normal(0, 3, 1000), np.array([2.5]), np.array([1.0, -3.0]),
             np.concatenate([np.full(50, 0.25), rng.normal(0.25, 1e-3, 49)]),
             np.concatenate([np.full(40, -1.5), np.full(40, 2.0)]),
             np.round(rng.normal(0, 2, 5000), 1), -np.abs(rng.normal(5, 1, 4096)) * 1e-300,
             rng.normal(0, 1, 300001) * 1e200,
             np.full(70000, 3.75), np.zeros(1000), np.concatenate([np.zeros(500), -np.zeros(500)]),
             np.concatenate([np.full(1000, 1.0), np.full(1000, np.nextafter(1.0, 2.0))]),      # the two middle elements differ in the last bit
             np.concatenate([np.full(1000, 1.0 - 2.0 ** -30), np.full(1000, 1.0)]),            # ... and lie in different 22-bit groups
             np.concatenate([rng.normal(0, 1, 999), [np.inf, -np.inf]]),
             rng.standard_cauchy(1000000)]
    for a in cases:
        got = float(images.median_device(torch.from_numpy(a).cuda()).cpu())
        assert got == np.median(a), (a.size, got, np.median(a))
    rows = rng.normal(1.0, 5.0, (7, 2048))
    got = images.median_device(torch.from_numpy(rows).cuda()).cpu().numpy()
    assert np.array_equal(got, np.median(rows, axis=1))
    a, b = torch.from_numpy(cases[0]).cuda(), torch.from_numpy(rng.normal(2, 1, 1001)).cuda()
    for u, v in ((a, b), (b, a)):
        mu, mv = images.median_device_pair(u, v)
        assert float(mu.cpu()) == np.median(u.cpu().numpy()) and float(mv.cpu()) == np.median(v.cpu().numpy())


def _batch_catalogue():
    sizes = [700, 5, 0, 2300, 64, 1500]            # incl. fewer particles than k, and an empty object
    parts = [synthetic.particles(max(n, 1), seed=70 + i, scale_kpc=0.6 + 0.3 * i) for i, n in enumerate(sizes)]
    cat = {k: np.concatenate([p[k][:n] for p, n in zip(parts, sizes)]) for k in ('X', 'Y', 'Masses')}
    off = np.concatenate([[0], np.cumsum(sizes)])
    rng = np.random.default_rng(4)
    cat['L'] = cat['Masses'] * (0.5 + rng.random(cat['Masses'].size))
    return cat, off, sizes


@pytest.mark.parametrize('smoothing', [False, ('convolved_gaussian', 0.25), ('adaptive', 8), ('adaptive', 12)])
def test_core_batch_equals_loop(smoothing):
    """core_batch (one pass over all objects: image id in the k-NN sort key and in the tile key) == a loop
    of core() over the objects == the oracle; NGP counts bit-exact"""
    from synthobs_b200.morph import images
    cat, off, sizes = _batch_catalogue()
    b = images.core_batch(cat['X'], cat['Y'], cat['L'], off, resolution=0.08, ndim=60, smoothing=smoothing)
    assert b.hist.shape == (len(sizes), 60, 60) and b.data.shape == (len(sizes), 60, 60)
    for g, n in enumerate(sizes):
        s, e = off[g], off[g + 1]
        if n == 0:
            assert not b.hist[g].any() and not b.data[g].any()
            continue
        one = images.core(cat['X'][s:e], cat['Y'][s:e], cat['L'][s:e], resolution=0.08, ndim=60, smoothing=smoothing)
        assert np.array_equal(b.hist[g], one.hist)
        np.testing.assert_allclose(b.simple[g], one.simple, rtol=1e-6)
        assert_image_close(b.data[g], one.data, tol=2e-6)
        ref = image_oracle.core(cat['X'][s:e], cat['Y'][s:e], cat['L'][s:e], resolution=0.08, ndim=60, smoothing=smoothing)
        assert np.array_equal(b.hist[g], ref['hist'])
        assert_image_close(b.data[g], ref['data'])


def test_observed_particle_batch_equals_loop():
    """observed.particle_batch (segmented medians, batched deposit, PSF FFT, rebin) == loop of observed.particle;
    multi-channel weights"""
    from synthobs_b200.morph import images
    images.filter_info.update({'JWST.NIRCAM.F150W': {'pixel_scale': 0.031}})
    cat, off, sizes = _batch_catalogue()
    cosmo = synthetic.FixedCosmology()
    L2 = np.stack([cat['L'], 0.3 * cat['L'][::-1].copy()])
    obs = images.observed('JWST.NIRCAM.F150W', cosmo, 8.0, 1.0, smoothing=('adaptive', 8.), PSF=synthetic.GaussianPSF(2.1),
                          super_sampling=2, xoffset_pix=0.2, yoffset_pix=-0.1)
    X0, Y0 = cat['X'].copy(), cat['Y'].copy()
    out = obs.particle_batch(cat['X'], cat['Y'], L2, off)
    assert np.array_equal(cat['X'], X0) and np.array_equal(cat['Y'], Y0)        # batch entry leaves the inputs alone
    assert out['data'].shape == (len(sizes), 2, obs.ndim, obs.ndim)
    for g, n in enumerate(sizes):
        if n == 0:
            assert not out['data'][g].any()
            continue
        s, e = off[g], off[g + 1]
        for c in range(2):
            one = obs.particle(cat['X'][s:e].copy(), cat['Y'][s:e].copy(), L2[c, s:e].copy())
            assert_image_close(out['no_PSF'][g, c], one.img.no_PSF, tol=2e-6)
            assert_image_close(out['data'][g, c], one.img.data, tol=2e-6)
    med = images.median_segmented_device(images.dv.to_dev(cat['X']), images.dv.to_dev(off, __import__('torch').int64)).cpu().numpy()
    for g, n in enumerate(sizes):
        if n:
            assert med[g] == np.median(cat['X'][off[g]:off[g + 1]])
        else:
            assert np.isnan(med[g])


@pytest.mark.parametrize('k', [1, 4, 8, 16, 30])
def test_knn_exact_ties_and_stacks_of_identical_positions(k):
    """the tile search (default for small calls) screens in FP32 and decides in float64: exact ties of the k-th
    distance (a lattice), more identical positions than a query's list keeps (the hand-over to the quadtree walk),
    points far outside the bulk and a nearly collinear streak must all come out equal to cKDTree (images.py:83-85)"""
    from synthobs_b200.morph import images
    from synthobs_b200 import device as dv
    rng = np.random.default_rng(7 + k)
    gx, gy = np.meshgrid(np.arange(40) * 0.05 - 1.0, np.arange(40) * 0.05 - 1.0)       # 1600 lattice points: ties everywhere
    stack = np.tile([[0.3141, -0.2718]], (60, 1))                                       # 60 identical positions
    pairs = np.repeat(rng.uniform(-2, 2, (50, 2)), 3, axis=0)                           # triples
    far = rng.uniform(-3.9, 3.9, (40, 2))
    streak = np.column_stack([np.linspace(-3, 3, 700), 1e-9 * rng.standard_normal(700) + 2.5])
    cloud = rng.normal(0, 0.7, (3000, 2))
    P = np.concatenate([np.column_stack([gx.ravel(), gy.ravel()]), stack, pairs, far, streak, cloud])
    P = P[rng.permutation(P.shape[0])]
    X, Y = np.ascontiguousarray(P[:, 0]), np.ascontiguousarray(P[:, 1])
    hw = 4.0
    width, G = images.pixel_grid(2 * hw / 100, 100)
    Xd, Yd = dv.to_dev(X), dv.to_dev(Y)
    ix, _ = images.ngp_index_device(Xd, Yd, dv.to_dev(G), 100, hw)
    dk = images.knn_device(Xd, Yd, ix, k, hw).cpu().numpy()
    sel = image_oracle.select(X, Y, 2 * hw)
    ref = image_oracle.kth_neighbour_distance(X[sel], Y[sel], k)
    assert np.array_equal(dk[~sel], np.zeros((~sel).sum()))
    np.testing.assert_allclose(dk[sel], ref, rtol=4e-16, atol=0)


@pytest.mark.parametrize('n', [3000, 50000])
def test_knn_batch_of_tiny_images_and_invalid_image_ids(n):
    """batch mode (image id per particle, config 4): many tiny images, some with fewer than k particles (dk = inf),
    equal to cKDTree per image for both searches (n picks the tile kernel / the quadtree walk); a particle whose image id
    lies outside [0, n_img) counts as outside its image (dk = 0) instead of corrupting the sort"""
    import torch
    from synthobs_b200.morph import images
    from synthobs_b200 import device as dv
    rng = np.random.default_rng(n)
    k, n_img, hw = 6, n // 12, 4.0
    X, Y = rng.normal(0, 0.8, n), rng.normal(0, 0.8, n)
    ids = np.sort(rng.integers(0, n_img, n)).astype(np.int32)
    bad = rng.choice(n, 25, replace=False)
    ids_in = ids.copy()
    ids_in[bad[:12]] = n_img + 7
    ids_in[bad[12:]] = -3
    width, G = images.pixel_grid(2 * hw / 100, 100)
    Xd, Yd = dv.to_dev(X), dv.to_dev(Y)
    ix, _ = images.ngp_index_device(Xd, Yd, dv.to_dev(G), 100, hw)
    dk = images.knn_device(Xd, Yd, ix, k, hw, img=dv.to_dev(ids_in, torch.int32), n_img=n_img).cpu().numpy()
    ok = image_oracle.select(X, Y, 2 * hw)
    ok[bad] = False
    ref = np.zeros(n)
    for g in np.unique(ids):
        m = ok & (ids == g)
        if m.sum() == 0:
            continue
        ref[m] = image_oracle.kth_neighbour_distance(X[m], Y[m], k) if m.sum() >= k else np.inf
    assert np.array_equal(np.isinf(dk), np.isinf(ref))
    f = np.isfinite(ref)
    np.testing.assert_allclose(dk[f], ref[f], rtol=4e-16, atol=0)


@pytest.mark.parametrize('k', [8, 12])
def test_knn_floor_shortcut_leaves_the_image_unchanged(k):
    """FWHM = max(dk, 0.01) (images.py:89): with floor = 0.01 the k-NN may return any upper bound <= 0.01 for
    particles that have k neighbours inside the floor radius.  Above the floor the distances are the exact ones,
    below they stay below, and the image is bit-identical to the one from exact distances."""
    import torch
    from synthobs_b200 import device as dv
    from synthobs_b200.morph import images
    p = synthetic.particles(60000, seed=5, scale_kpc=0.4)            # dense: many particles below the floor
    Xd, Yd, Ld = dv.to_dev(p['X']), dv.to_dev(p['Y']), dv.to_dev(p['Masses']).reshape(1, -1)
    res, ndim = 0.02, 128
    width, G = images.pixel_grid(res, ndim)
    ix, _ = images.ngp_index_device(Xd, Yd, dv.to_dev(G), ndim, width / 2.)
    exact = images.knn_device(Xd, Yd, ix, k, width / 2.)
    fl = images.knn_device(Xd, Yd, ix, k, width / 2., floor=images.FWHM_FLOOR)
    above = exact > images.FWHM_FLOOR
    assert 0.05 < float((~above).double().mean()) < 0.95             # the case exercises both branches
    assert bool((fl[above] == exact[above]).all())
    assert bool((fl[~above] <= images.FWHM_FLOOR).all()) and bool((fl[~above] >= exact[~above]).all())
    a = images.deposit_device(Xd, Yd, Ld, res, ndim, adaptive_k=k, want_ngp=False, exact_dk=True)['data']
    b = images.deposit_device(Xd, Yd, Ld, res, ndim, adaptive_k=k, want_ngp=False)['data']
    assert bool((a == b).all())
