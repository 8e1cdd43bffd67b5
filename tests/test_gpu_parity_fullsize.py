"""GPU parity AT THE BENCHMARKED SIZES: the CUDA path (through the C-ABI) against the float64 ORACLE -- not
against itself -- at BASELINE.json's configurations:

  config[1]  1e5 particles, 51 x 13 x 9000 grid, ISM + BC dust, 10 filters (reference: synthobs/sed/models.py:88-135,
             139-200, 229-314): grid indices bit-exact, per-particle values and the five spectra 1e-5, band sums 1e-4;
  config[2]  1e6 particles -> 1024^2 super-sampled (512^2, ss 2), ('adaptive', 8), PSF FFT, rebin
             (synthobs/morph/images.py:83-97, 181-213): images 1e-5 (peak and L2), flux 2e-6;
  config[3]  20 ragged galaxies of 1e3..1e5 particles in ONE batched pass vs a loop of oracle calls;
  config[4]  a 3e6-particle, 8192^2 image with two dense clumps (tens of thousands of FP32 contributions per pixel)
             vs the oracle on 512^2 windows, for one particle share and for three shares added.

The oracle reaches these sizes because (i) its spectra are summed over particle blocks on all host cores (NumPy
releases the GIL) and (ii) image_oracle.adaptive_window skips the terms that are exactly 0.0 in float64 (pinned
against adaptive_separable / the golden images in tests/test_oracle_golden.py)."""

import contextlib
import io
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from oracle import image_oracle, sed_oracle
from synthobs_b200 import providers, synthetic

pytestmark = pytest.mark.gpu

RTOL_PARTICLE = 1e-5       # north_star: images and SEDs within 1e-5 relative (FP32)
RTOL_SUM = 1e-4            # north_star: filter fluxes within 1e-4
Z = 8.0
SPECTRA = ('stellar', 'intrinsic', 'BC', 'total', 'no_HII_no_BC')
SIMPLE = ('simple', {'slope': -1.0})


def assert_image_close(a, b, tol=1e-5):
    """the image tolerances of tests/test_gpu_images.py: 1e-5 of the peak per pixel and 1e-5 relative in L2"""
    peak = np.abs(b).max()
    assert a.shape == b.shape
    assert np.abs(a - b).max() <= tol * peak, (np.abs(a - b).max() / peak)
    assert np.linalg.norm(a - b) <= tol * np.linalg.norm(b), np.linalg.norm(a - b) / np.linalg.norm(b)


def oracle_spectra(grid, p, sl, kI, kB, fesc, log10t_BC=7.0, block=2500):
    """sed_oracle.sed_spectra of the particles p[sl], summed over particle blocks evaluated on all host cores"""
    idx = np.arange(p['Masses'].size)[sl]
    blocks = [idx[s:s + block] for s in range(0, idx.size, block)]

    def one(q):
        return sed_oracle.sed_spectra(grid, p['Masses'][q], p['Ages'][q], p['Metallicities'][q],
                                      tauVs_ISM=p['tauVs_ISM'][q], tauVs_BC=p['tauVs_BC'][q], k_ISM_lam=kI, k_BC_lam=kB,
                                      fesc=fesc, log10t_BC=log10t_BC)
    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        parts = list(ex.map(one, blocks))
    return {k: np.sum([r[k] for r in parts], axis=0) for k in SPECTRA}


@pytest.fixture(scope='module')
def full_model():
    """the bench model (bench.py setup_model): 51 x 13 x 9000 grid, 10 observed-frame filters at z = 8, plus the
    oracle's own float64 filter tables (sed_oracle.fnu_grid restates models.py:53-67)"""
    from synthobs_b200.sed import models
    grid = synthetic.ssp_grid(n_lam=9000, seed=0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = models.define_model(grid)
    m.dust_ISM = SIMPLE
    m.dust_BC = SIMPLE
    F = synthetic.filter_set(m.lam * (1. + Z), n_filters=10, kind='gauss', lo=12000., hi=50000., prefix='JWST.FAKE.')
    cosmo = synthetic.FixedCosmology()
    m.create_Fnu_grid(F, Z, cosmo)
    ogrid = sed_oracle.ensure_incident(synthetic.ssp_grid(n_lam=9000, seed=0))
    otabs = sed_oracle.fnu_grid(ogrid, F, Z, cosmo.dl_cm, providers.madau)
    k = {f: float(providers.simple({'slope': -1.0}).tau(F[f].pivwv() / (1. + Z))) for f in F['filters']}
    return m, F, ogrid, otabs, k


def test_config2_filter_tables_vs_oracle(full_model):
    """create_Fnu_grid at 9000 wavelengths (the K = 9000 tcgen05 contraction) vs the float64 trapz tables"""
    model, F, ogrid, otabs, k = full_model
    for f in F['filters']:
        for c in sed_oracle.COMPONENTS:
            ref = otabs[f][c]
            np.testing.assert_allclose(getattr(model.Fnu[f], c), ref, rtol=RTOL_PARTICLE, atol=1e-7 * np.abs(ref).max())


def test_config2_broadband_vs_oracle(full_model):
    """generate_Fnu / generate_Fnu_array / grid indices for 1e5 particles, ISM + BC dust, 10 filters"""
    from synthobs_b200.sed import models
    model, F, ogrid, otabs, k = full_model
    p = synthetic.particles(100000, seed=41, lognormal_mass=True)
    args = (p['Masses'], p['Ages'], p['Metallicities'])
    kw = dict(tauVs_ISM=p['tauVs_ISM'], tauVs_BC=p['tauVs_BC'], fesc=0.1)
    oargs = (ogrid['log10age'], ogrid['log10Z']) + args
    # bit-exact grid indices (models.py:108-114, 124)
    ia, iZ, young = models.ngp_cells(model, p['Ages'], p['Metallicities'])
    ra, rz, la = sed_oracle.ngp_indices(ogrid['log10age'], ogrid['log10Z'], p['Ages'], p['Metallicities'])
    assert np.array_equal(ia, ra) and np.array_equal(iZ, rz) and np.array_equal(young, la < 7.0)
    # band sums against the oracle's OWN tables (tables + particle pass together: 1e-4)
    got = models.generate_Fnu(model, F, *args, **kw)
    ref = sed_oracle.broadband(otabs, F['filters'], *oargs, k_ISM=k, k_BC=k, **kw)
    np.testing.assert_allclose([got[f] for f in F['filters']], [ref[f] for f in F['filters']], rtol=RTOL_SUM)
    # per-particle arrays, every filter (one pass), against the oracle on the tables the kernel was given: 1e-5
    arrs = models.generate_Fnu_arrays(model, F, *args, **kw)
    for f in F['filters']:
        tab = {c: getattr(model.Fnu[f], c) for c in sed_oracle.COMPONENTS}
        r = sed_oracle.broadband_array(tab, *oargs, k_ISM=k[f], k_BC=k[f], **kw)
        np.testing.assert_allclose(arrs[f], r, rtol=RTOL_PARTICLE, atol=1e-30, err_msg=f)
        assert arrs[f].sum() == pytest.approx(got[f], rel=1e-6)
    # the single-filter reference call (models.py:153)
    f = F['filters'][3]
    one = models.generate_Fnu_array(model, F, f, *args, **kw)
    assert np.array_equal(one, arrs[f])


def test_config2_spectra_vs_oracle(full_model):
    """generate_SED: the five spectra over 9000 wavelengths for 1e5 particles (histogram -> tcgen05 GEMM with K = 1326,
    per-particle dust kernel) vs sed_oracle.sed_spectra (models.py:265-310)"""
    from synthobs_b200.sed import models
    model, F, ogrid, otabs, k = full_model
    p = synthetic.particles(100000, seed=42, lognormal_mass=True)
    with contextlib.redirect_stdout(io.StringIO()):
        o = models.generate_SED(model, p['Masses'], p['Ages'], p['Metallicities'], tauVs_ISM=p['tauVs_ISM'],
                                tauVs_BC=p['tauVs_BC'], fesc=0.1)
    kl = providers.simple({'slope': -1.0}).tau(ogrid['lam'])
    r = oracle_spectra(ogrid, p, slice(None), kl, kl, 0.1)
    for key in SPECTRA:
        got = getattr(o, key).lnu
        assert got.shape == (9000,)
        np.testing.assert_allclose(got, r[key], rtol=RTOL_PARTICLE, atol=1e-6 * np.abs(r[key]).max(), err_msg=key)
    with np.errstate(divide='ignore', invalid='ignore'):
        fesc_ref = r['total'] / r['intrinsic']
    ok = np.isfinite(fesc_ref) & (r['intrinsic'] > 1e-6 * r['intrinsic'].max())
    np.testing.assert_allclose(o.f_esc[ok], fesc_ref[ok], rtol=1e-4)


def test_config3_observed_image_vs_oracle():
    """1e6 particles, 512 x 512 observed frame, super-sampling 2 (1024^2), ('adaptive', 8), Gaussian PSF, rebin:
    two channels through the multi-channel pipeline the bench times, against the oracle's observed.particle
    (images.py:181-213).  Pixel indices and counts bit-exact; images 1e-5."""
    import torch
    from synthobs_b200.morph import images
    n = 1000000
    p = synthetic.particles(n, seed=200, scale_kpc=3.0)            # the bench catalogue (bench.py run_imaging)
    filters = ['Q.A', 'Q.B']
    for f in filters:
        images.filter_info[f] = {'pixel_scale': 0.031}
    cosmo = synthetic.FixedCosmology()
    width_arcsec = 512 * 0.031 + 1e-9
    psfs = [synthetic.GaussianPSF(1.5), synthetic.GaussianPSF(2.9)]
    obs = [images.observed(f, cosmo, Z, width_arcsec, smoothing=('adaptive', 8.), PSF=psf, super_sampling=2)
           for f, psf in zip(filters, psfs)]
    assert obs[0].ndim == 512 and obs[0].ndim_super == 1024
    L = np.stack([p['Masses'] * (0.5 + np.random.default_rng(7).random(n)), p['Masses']])
    Xd = torch.from_numpy(p['X']).cuda()
    Yd = torch.from_numpy(p['Y']).cuda()
    mx, my = images.median_device_pair(Xd, Yd)
    assert float(mx) == np.median(p['X']) and float(my) == np.median(p['Y'])      # np.median semantics, bit-exact
    r = images.particle_device_multi(obs, Xd - mx, Yd - my, torch.from_numpy(L).cuda())
    # oracle
    g = image_oracle.observed_geometry(0.031, cosmo.arcsec_per_kpc, width_arcsec, super_sampling=2)
    X = p['X'] - np.median(p['X'])
    Y = p['Y'] - np.median(p['Y'])
    width, G = image_oracle.pixel_grid(g['resolution'], g['ndim_super'])
    sel = image_oracle.select(X, Y, width)
    assert r['stats']['selected'] == sel.sum()
    dk = image_oracle.kth_neighbour_distance(X[sel], Y[sel], 8)
    sup = image_oracle.adaptive_tiled(G, X[sel], Y[sel], L[:, sel], dk)
    half = g['width_arcsec'] / g['native_pixel_scale'] / 2.
    xx = np.linspace(-half, half, g['ndim_super'])
    for c in range(2):
        got_sup = r['super_no_PSF'][c].double().cpu().numpy()
        assert_image_close(got_sup, sup[c])
        assert got_sup.sum() == pytest.approx(L[c][sel].sum(), rel=2e-6)
        conv = image_oracle.convolve_fft_same(sup[c], psfs[c].f(xx, xx))
        assert_image_close(r['super_data'][c].double().cpu().numpy(), conv)
        assert_image_close(r['no_PSF'][c].double().cpu().numpy(), image_oracle.rebin(sup[c], (512, 512)))
        assert_image_close(r['data'][c].double().cpu().numpy(), image_oracle.rebin(conv, (512, 512)))
    # the reference-named call on the same catalogue: observed.particle (in-place recentring) == the multi-channel path
    Xc, Yc = p['X'].copy(), p['Y'].copy()
    one = obs[1].particle(Xc, Yc, L[1])
    np.testing.assert_array_equal(Xc, X)
    assert_image_close(one.img.data, image_oracle.rebin(conv, (512, 512)))
    # NGP products at this size: counts and pixel indices bit-exact against the literal argmin (images.py:66-69)
    h = images.deposit_device(Xd - mx, Yd - my, torch.from_numpy(L[1:2]).cuda(), obs[0].resolution, 1024, adaptive_k=None)
    i = image_oracle.ngp_pixels(G, X[sel])
    j = image_oracle.ngp_pixels(G, Y[sel])
    assert np.array_equal(h['ix'].cpu().numpy()[sel], i) and np.array_equal(h['iy'].cpu().numpy()[sel], j)
    hist = np.zeros((1024, 1024))
    np.add.at(hist, (j, i), 1.0)
    assert np.array_equal(h['hist'].cpu().numpy(), hist)


def test_config4_ragged_batch_vs_oracle_loop(full_model):
    """20 galaxies of 1e3..1e5 particles (log-uniform, the FLARES-style ragged catalogue) in ONE batched pass:
    band fluxes, spectra and adaptive images vs a loop of oracle calls over the galaxies
    (examples/core_manytestobjects.py:17-22 is that loop upstream)"""
    from synthobs_b200.morph import images
    from synthobs_b200.sed import models
    model, F, ogrid, otabs, k = full_model
    b = synthetic.galaxy_batch(20, n_min=1000, n_max=100000, seed=5)
    off = b['offsets']
    sizes = np.diff(off)
    assert sizes.min() < 3000 and sizes.max() > 50000
    args = (b['Masses'], b['Ages'], b['Metallicities'])
    kw = dict(tauVs_ISM=b['tauVs_ISM'], tauVs_BC=b['tauVs_BC'], fesc=0.05)
    rows = models.generate_Fnu_batch(model, F, *args, off, **kw)
    for gi in range(20):
        q = slice(off[gi], off[gi + 1])
        ref = sed_oracle.broadband(otabs, F['filters'], ogrid['log10age'], ogrid['log10Z'], b['Masses'][q], b['Ages'][q],
                                   b['Metallicities'][q], tauVs_ISM=b['tauVs_ISM'][q], tauVs_BC=b['tauVs_BC'][q],
                                   k_ISM=k, k_BC=k, fesc=0.05)
        np.testing.assert_allclose(rows[gi], [ref[f] for f in F['filters']], rtol=RTOL_SUM)
    # spectra: whole batch on the GPU, the smallest, the largest and two more galaxies checked against the oracle
    out = models.generate_SED_batch(model, *args, off, **kw)
    kl = providers.simple({'slope': -1.0}).tau(ogrid['lam'])
    order = np.argsort(sizes)
    for gi in (order[0], order[-1], order[7], order[13]):
        r = oracle_spectra(ogrid, b, slice(off[gi], off[gi + 1]), kl, kl, 0.05)
        for key in SPECTRA:
            np.testing.assert_allclose(out[key][gi], r[key], rtol=RTOL_PARTICLE, atol=1e-6 * np.abs(r[key]).max(),
                                       err_msg='%s galaxy %d' % (key, gi))
    # images: 100 x 100 physical, ('adaptive', 16), all galaxies in one pass
    L = b['Masses'] * (0.5 + np.random.default_rng(3).random(b['Masses'].size))
    img = images.core_batch(b['X'], b['Y'], L, off, resolution=0.1, ndim=100, smoothing=('adaptive', 16))
    for gi in range(20):
        q = slice(off[gi], off[gi + 1])
        o = image_oracle.core(b['X'][q], b['Y'][q], L[q], resolution=0.1, ndim=100, smoothing=('adaptive', 16))
        assert np.array_equal(img.hist[gi], o['hist'])
        np.testing.assert_allclose(img.simple[gi], o['simple'], rtol=2e-7)
        assert_image_close(img.data[gi], o['data'])


def test_config5_huge_image_windows_vs_oracle():
    """3e6 particles onto 8192 x 8192 with a moderate and an extreme clump (the extreme one stacks tens of thousands of
    FP32 contributions per pixel at the FWHM floor): one particle share and three spatial shares added, each
    against the float64 oracle on 512 x 512 windows around the clumps and in the field."""
    import torch
    from synthobs_b200 import device as dv
    from synthobs_b200.morph import images
    N, ndim, S = 3000000, 8192, 10.0
    rng = np.random.default_rng(99)
    X = S * rng.standard_normal(N)
    Y = S * rng.standard_normal(N)
    m = N // 20
    X[:m] = 3.0 + 0.5 * rng.standard_normal(m)
    Y[:m] = -7.0 + 0.5 * rng.standard_normal(m)
    X[m:2 * m] = -11.0 + 0.03 * rng.standard_normal(m)
    Y[m:2 * m] = 5.0 + 0.03 * rng.standard_normal(m)
    L = 8.47e5 * (0.5 + rng.random(N))
    res = 8.0 * S / ndim
    width, G = image_oracle.pixel_grid(res, ndim)
    sel = image_oracle.select(X, Y, width)
    Xs, Ys, Ls = X[sel], Y[sel], L[sel]
    dk = image_oracle.kth_neighbour_distance(Xs, Ys, 8)
    Xd, Yd, Ld = dv.to_dev(X), dv.to_dev(Y), dv.to_dev(L)
    ix, iy = images.ngp_index_device(Xd, Yd, dv.to_dev(G), ndim, width / 2.)
    assert int((ix >= 0).sum()) == sel.sum()
    # exact k-NN at this size (floor shortcut off): equal to cKDTree to rounding
    got_dk = images.knn_device(Xd, Yd, ix, 8, width / 2.).cpu().numpy()[sel]
    np.testing.assert_allclose(got_dk, dk, rtol=4e-16, atol=0)
    from synthobs_b200 import dist as sdist
    whole = images.deposit_device(Xd, Yd, Ld.reshape(1, -1), res, ndim, adaptive_k=8, want_ngp=False)['data'][0]
    # the multi-GPU path on one GPU: the partial images of the 3 shares (what 3 ranks would deposit), added
    shares = torch.zeros_like(whole)
    for part in range(3):
        shares += sdist.image_huge(Xd, Yd, Ld, res, ndim, 8, rank=part, n_ranks=3, reduce=False)
    flux = Ls.sum()
    assert float(whole.double().sum()) == pytest.approx(flux, rel=2e-6)
    assert float(shares.double().sum()) == pytest.approx(flux, rel=2e-6)

    def centre(x):
        return int(np.clip(np.abs(G - x).argmin() - 256, 0, ndim - 512))
    for cx, cy in ((-11.0, 5.0), (3.0, -7.0), (0.0, 0.0), (-39.0, 39.0)):
        c0, r0 = centre(cx), centre(cy)
        ref = image_oracle.adaptive_window(G, Xs, Ys, Ls, dk, (r0, r0 + 512), (c0, c0 + 512))
        assert ref.max() > 0
        for img in (whole, shares):
            assert_image_close(img[r0:r0 + 512, c0:c0 + 512].double().cpu().numpy(), ref)
