"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/synthobs, loaded through oracle/ref_loader.py) on synthetic
inputs.  Run in the build container only (the reference cannot travel to the
GPU box):

    python tests/golden/make_golden.py

Each fixture stores the inputs (when small), the seeds/shape parameters needed
to regenerate the SSP grid with synthobs_b200.synthetic, a checksum of that
grid, and the reference's outputs.  The fixtures pin oracle/ (CPU tests) and,
through it and directly, the CUDA path (-m gpu tests).
"""

import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, '..', '..')))

from oracle import ref_loader  # noqa: E402
from synthobs_b200 import synthetic  # noqa: E402

warnings.simplefilter('ignore')
GRID_KW = dict(n_lam=300, n_age=51, seed=0)


def grid_checksum(grid):
    return np.array([grid['stellar'].sum(), grid['nebular'].sum(), grid['log10Q'].sum(),
                     grid['stellar'][7, 3, 100], grid['nebular'][40, 11, 250]])


def make_model(ref, with_incident=False):
    grid = synthetic.ssp_grid(with_incident=with_incident, **GRID_KW)
    tmp = tempfile.mkdtemp()
    name = synthetic.write_grid_pickle(grid, tmp)
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        model = ref.sed.models.define_model(name, path_to_SPS_grid=tmp + '/')
    return model, grid


def sed_case(ref, tag, n, seed, dust_ISM, dust_BC, fesc, log10t_BC, observed, with_incident=False,
             pass_tau=True, lognormal_mass=True):
    models = ref.sed.models
    model, grid = make_model(ref, with_incident)
    model.dust_ISM = dust_ISM
    model.dust_BC = dust_BC
    p = synthetic.particles(n, seed=seed, lognormal_mass=lognormal_mass)
    z = 8.0
    cosmo = synthetic.FixedCosmology()
    if observed:
        F = synthetic.filter_set(model.lam * (1. + z), n_filters=4, kind='gauss', lo=12000., hi=50000.,
                                 prefix='JWST.FAKE.')
        model.create_Fnu_grid(F, z, cosmo)
        tabs, gen_arr, gen_sum = model.Fnu, models.generate_Fnu_array, models.generate_Fnu
    else:
        F = synthetic.filter_set(model.lam, n_filters=4, kind='tophat', lo=1500., hi=8000.)
        model.create_Lnu_grid(F)
        tabs, gen_arr, gen_sum = model.Lnu, models.generate_Lnu_array, models.generate_Lnu
    kw = dict(fesc=fesc, log10t_BC=log10t_BC)
    if pass_tau:
        kw.update(tauVs_ISM=p['tauVs_ISM'], tauVs_BC=p['tauVs_BC'])
    out = {'grid_checksum': grid_checksum(grid), 'with_incident': with_incident, 'n': n, 'seed': seed,
           'fesc': fesc, 'log10t_BC': log10t_BC, 'observed': observed, 'z': z, 'pass_tau': pass_tau,
           'dust_ISM': repr(dust_ISM), 'dust_BC': repr(dust_BC), 'dl_cm': cosmo.dl_cm}
    for k in ('Masses', 'Ages', 'Metallicities', 'tauVs_ISM', 'tauVs_BC'):
        out[k] = p[k]
    for fi, f in enumerate(F['filters']):
        for c in ('stellar_incident', 'stellar_transmitted', 'nebular'):
            out['tab_%d_%s' % (fi, c)] = getattr(tabs[f], c)
        out['arr_%d' % fi] = gen_arr(model, F, f, p['Masses'], p['Ages'], p['Metallicities'], **kw)
    tot = gen_sum(model, F, p['Masses'], p['Ages'], p['Metallicities'], **kw)
    out['sums'] = np.array([tot[f] for f in F['filters']])
    out['log10Q'] = models.generate_log10Q(model, p['Masses'], p['Ages'], p['Metallicities'])
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        o = models.generate_SED(model, p['Masses'], p['Ages'], p['Metallicities'], **kw)
    for k in ('stellar', 'intrinsic', 'BC', 'total', 'no_HII_no_BC'):
        out['sed_' + k] = getattr(o, k).lnu
    for k in ('f_esc', 'tau', 'tau_relV'):
        out['sed_' + k] = getattr(o, k)
    np.savez_compressed(os.path.join(HERE, 'sed_%s.npz' % tag), **out)
    print('sed_%s' % tag, 'sums', out['sums'])


def lines_case(ref, tag, n, seed, dust_ISM, dust_BC, fesc, log10t_BC, with_incident=False, pass_tau=True):
    """EmissionLines.get_line_luminosity / get_line_luminosities of the unmodified reference
    (models.py:326-448) on a synthetic lines.p grid, single and blended lines."""
    import contextlib
    import io
    models = ref.sed.models
    lgrid = synthetic.lines_grid(n_age=GRID_KW['n_age'], seed=0, with_incident=with_incident)
    tmp = tempfile.mkdtemp()
    name = synthetic.write_lines_pickle(lgrid, tmp)
    with contextlib.redirect_stdout(io.StringIO()):
        el = models.EmissionLines(name, dust_ISM=dust_ISM, dust_BC=dust_BC, verbose=False, path_to_SPS_grid=tmp)
    p = synthetic.particles(n, seed=seed, lognormal_mass=True)
    kw = dict(fesc=fesc, log10t_BC=log10t_BC)
    if pass_tau:
        kw.update(tauVs_ISM=p['tauVs_ISM'], tauVs_BC=p['tauVs_BC'])
    lines = ['HI6563', 'OIII4959,OIII5007', 'HI4861', 'OII3726,OII3729', 'CIII1907,CIII1909']
    with contextlib.redirect_stdout(io.StringIO()):
        o = el.get_line_luminosities(lines, p['Masses'], p['Ages'], p['Metallicities'], **kw)
        single = el.get_line_luminosity('NII6583', p['Masses'], p['Ages'], p['Metallicities'], **kw)
    out = {'n': n, 'seed': seed, 'fesc': fesc, 'log10t_BC': log10t_BC, 'with_incident': with_incident,
           'pass_tau': pass_tau, 'dust_ISM': repr(dust_ISM), 'dust_BC': repr(dust_BC), 'lines': np.array(lines),
           'grid_checksum': np.array([lgrid['HI6563']['luminosity'].sum(), lgrid['OII3729']['stellar_continuum'].sum()])}
    for k in ('Masses', 'Ages', 'Metallicities', 'tauVs_ISM', 'tauVs_BC'):
        out[k] = p[k]
    for q in ('luminosity', 'continuum', 'EW'):
        out[q] = np.array([o[l][q] for l in lines])
        out['single_' + q] = single[q]
    np.savez_compressed(os.path.join(HERE, 'lines_%s.npz' % tag), **out)
    print('lines_%s' % tag, 'luminosity', out['luminosity'], 'EW', out['EW'])


def sed_index_case(ref):
    """NGP indices as the reference sees them, including adversarial samples at
    cell midpoints +- ulps: the per-filter table is replaced by a cell-id code
    so generate_Lnu_array returns the (ia, iZ) it looked up (models.py:113-114,
    :131-133)."""
    models = ref.sed.models
    model, grid = make_model(ref)
    na, nZ = grid['log10age'].size, grid['log10Z'].size
    code = (np.arange(na)[:, None] * 100 + np.arange(nZ)[None, :]).astype(np.float64)
    F = {'filters': ['code']}
    model.Lnu = {'code': type('e', (), {})()}
    model.Lnu['code'].stellar_incident = code
    model.Lnu['code'].stellar_transmitted = code * 0 + 100000.0   # marks "young" lookups
    model.Lnu['code'].nebular = code * 0
    ages_adv = synthetic.adversarial_ages(grid['log10age'], ulps=(0, 1, 2, 3))
    zmid = 0.5 * (grid['log10Z'][:-1] + grid['log10Z'][1:])
    z_adv = 10.0 ** zmid
    z_adv = np.concatenate([z_adv, np.nextafter(z_adv, 1), np.nextafter(z_adv, 0)])
    rng = np.random.default_rng(5)
    ages_rand = 10 ** rng.uniform(-1.0, 5.0, 3000)
    z_rand = 10 ** rng.uniform(-6, -1, 3000)
    # young/old threshold samples: log10(Age)+6 around 7.0
    a_thr = np.array([10.0, np.nextafter(10.0, 0), np.nextafter(10.0, 100), 9.999999999, 10.000000001])
    special_age = np.array([0.0, np.inf, 1e-300, 1e300, 5e-324])
    special_z = np.array([0.0, np.inf, 1e-300, 1.0, 5e-324])
    Ages = np.concatenate([ages_adv, np.resize(ages_rand, z_adv.size), ages_rand, a_thr, special_age])
    Zs = np.concatenate([np.resize(z_rand, ages_adv.size), z_adv, z_rand, np.resize(z_rand, a_thr.size), special_z])
    M = np.ones(Ages.size)
    with np.errstate(all='ignore'):
        old = models.generate_Lnu_array(model, F, 'code', M, Ages, Zs, fesc=1.0)   # young: 1*inc, old: inc
        yng = models.generate_Lnu_array(model, F, 'code', M, Ages, Zs, fesc=0.0)   # young: 100000, old: inc
    ia = (old // 100).astype(np.int64)
    iZ = (old % 100).astype(np.int64)
    young = yng == 100000.0
    np.savez_compressed(os.path.join(HERE, 'sed_indices.npz'), Ages=Ages, Metallicities=Zs, ia=ia, iZ=iZ,
                        young=young, grid_checksum=grid_checksum(grid))
    print('sed_indices', Ages.size, 'young', young.sum())


def image_case(ref, tag, n, seed, resolution, ndim, smoothing, scale_kpc=1.5, extra=None):
    images = ref.morph.images
    p = synthetic.particles(n, seed=seed, scale_kpc=scale_kpc)
    X, Y, L = p['X'], p['Y'], p['Masses'] * (0.5 + np.random.default_rng(seed).random(n))
    if extra is not None:
        X, Y, L = extra(X, Y, L)
    with np.errstate(all='ignore'):
        img = images.core(X.copy(), Y.copy(), L.copy(), resolution=resolution, ndim=ndim, smoothing=smoothing)
    sm = np.array(['none', 0.0]) if not smoothing else np.array([smoothing[0], float(smoothing[1])])
    np.savez_compressed(os.path.join(HERE, 'img_%s.npz' % tag), X=X, Y=Y, L=L, resolution=resolution, ndim=ndim,
                        smoothing=sm, hist=img.hist, simple=img.simple, data=img.data)
    print('img_%s' % tag, 'sum', img.data.sum(), 'Lsum', L.sum())


def pixel_index_case(ref):
    """NGP pixel indices as the reference sees them (images.py:67), one
    particle per call, at pixel midpoints +- ulps and at the image edge."""
    images = ref.morph.images
    resolution, ndim = 0.1, 50
    width = ndim * resolution
    G = np.linspace(-width / 2., width / 2., ndim)
    mids = 0.5 * (G[:-1] + G[1:])
    xs = [mids]
    for u in range(1, 3):
        up, dn = mids.copy(), mids.copy()
        for _ in range(u):
            up, dn = np.nextafter(up, 10), np.nextafter(dn, -10)
        xs += [up, dn]
    xs.append(np.array([np.nextafter(-width / 2, 0), np.nextafter(width / 2, 0), 0.0, -0.0, G[7], G[31]]))
    xs = np.concatenate(xs)
    idx = np.empty(xs.size, dtype=np.int64)
    for q, x in enumerate(xs):
        img = images.core(np.array([x]), np.array([0.123]), np.array([1.0]), resolution=resolution, ndim=ndim)
        idx[q] = int(np.argmax(img.hist.sum(axis=0)))
    np.savez_compressed(os.path.join(HERE, 'img_indices.npz'), x=xs, idx=idx, resolution=resolution, ndim=ndim)
    print('img_indices', xs.size)


def observed_case(ref):
    images = ref.morph.images
    cosmo = synthetic.FixedCosmology()
    z = 8.0
    n = 1500
    p = synthetic.particles(n, seed=11, scale_kpc=0.8)
    filters = ['JWST.NIRCAM.F150W', 'JWST.NIRCAM.F356W']
    PSFs = {'JWST.NIRCAM.F150W': synthetic.GaussianPSF(2.1), 'JWST.NIRCAM.F356W': synthetic.GaussianPSF(1.8)}
    rng = np.random.default_rng(3)
    L = {f: p['Masses'] * rng.random(n) for f in filters}
    out = {'X': p['X'].copy(), 'Y': p['Y'].copy(), 'z': z, 'width_arcsec': 1.2, 'super_sampling': 2,
           'fwhm': np.array([2.1, 1.8]), 'arcsec_per_kpc': cosmo.arcsec_per_kpc}
    for fi, f in enumerate(filters):
        out['L_%d' % fi] = L[f]
    # single-filter class API (images.py:114,181), keeps the super-sampled products
    X1, Y1 = p['X'].copy(), p['Y'].copy()
    with np.errstate(all='ignore'):
        o = images.observed(filters[0], cosmo, z, 1.2, smoothing=('adaptive', 8.), PSF=PSFs[filters[0]],
                            super_sampling=2, xoffset_pix=0.25, yoffset_pix=-0.4).particle(X1, Y1, L[filters[0]])
    out.update(obs_super_no_PSF=o.super.no_PSF, obs_super_data=o.super.data, obs_no_PSF=o.img.no_PSF,
               obs_data=o.img.data, obs_X_after=X1, obs_Y_after=Y1,
               obs_pixel_scale=o.img.pixel_scale, obs_pixel_scale_kpc=o.img.pixel_scale_kpc)
    # multi-filter function API (images.py:263); X, Y are mutated cumulatively per filter
    X2, Y2 = p['X'].copy(), p['Y'].copy()
    with np.errstate(all='ignore'):
        IMGs = images.particle(X2, Y2, L, filters, cosmo, z, 1.2, smoothing=('adaptive', 8.), PSFs=PSFs,
                               super_sampling=2)
    for fi, f in enumerate(filters):
        out['multi_%d_data' % fi] = IMGs[f].data
        out['multi_%d_no_PSF' % fi] = IMGs[f].no_PSF
    out['multi_X_after'] = X2
    np.savez_compressed(os.path.join(HERE, 'img_observed.npz'), **out)
    print('img_observed', out['obs_data'].shape, [out['multi_%d_data' % i].shape for i in range(2)])


def analytic_case(ref):
    """observed.Sersic, Sersic(), point() of the unmodified reference (images.py:219-257, 290-313,
    320-414).  Sersic2D is the restated astropy model (providers.Sersic2D, via the stub)."""
    images = ref.morph.images
    cosmo = synthetic.FixedCosmology()
    z = 8.0
    out = {}
    p = {'r_eff': 0.9, 'n': 1.7, 'ellip': 0.35, 'theta': 0.6}
    out['sersic_p'] = np.array([p['r_eff'], p['n'], p['ellip'], p['theta']])
    o = images.observed('JWST.NIRCAM.F150W', cosmo, z, 0.8, resampling_factor=2, PSF=synthetic.GaussianPSF(2.1),
                        super_sampling=3, xoffset_pix=0.3, yoffset_pix=-0.2).Sersic(3.5e28, p)
    out.update(ser_super_no_PSF=o.super.no_PSF, ser_super_data=o.super.data, ser_no_PSF=o.img.no_PSF,
               ser_data=o.img.data, ser_pixel_scale=o.img.pixel_scale, ser_pixel_scale_kpc=o.img.pixel_scale_kpc)
    filters = ['JWST.NIRCAM.F150W', 'JWST.NIRCAM.F356W']
    PSFs = {'JWST.NIRCAM.F150W': synthetic.GaussianPSF(2.1), 'JWST.NIRCAM.F356W': synthetic.GaussianPSF(1.8)}
    L = {'JWST.NIRCAM.F150W': 1.0e28, 'JWST.NIRCAM.F356W': 2.5e28}
    IMGs = images.Sersic(L, {'r_eff': 0.5, 'n': 4.0, 'ellip': 0.0, 'theta': 0.0}, filters, cosmo, z, 1.2, PSFs=PSFs,
                         super_sampling=2)
    for fi, f in enumerate(filters):
        out['sermulti_%d_data' % fi] = IMGs[f].data
        out['sermulti_%d_no_PSF' % fi] = IMGs[f].no_PSF
    # point source through a tabulated PSF (the protocol of the FITS-backed PSFs: .f, .data, .ndim, .width)
    psf = synthetic.airy_like_psf(ndim=101, sub=5, fwhm_pix=2.0, pixel_scale=0.031)
    o = images.point(7.0e3, 'JWST.NIRCAM.F150W', 0.55, PSF=psf, super_sampling=5, xoffset_pix=0.21, yoffset_pix=-0.37)
    out.update(pt_super_hist=o.super.hist, pt_super_simple=o.super.simple, pt_super_data=o.super.data,
               pt_no_PSF=o.img.no_PSF, pt_data=o.img.data, pt_ndim=o.img.ndim, pt_pixel_scale=o.img.pixel_scale)
    o = images.point(2.0, 'JWST.NIRCAM.F356W', 1.5, resampling_factor=2, super_sampling=3,
                     PSF=synthetic.airy_like_psf(ndim=75, sub=5, fwhm_pix=1.4, pixel_scale=0.063, seed=4))
    out.update(pt2_super_data=o.super.data, pt2_no_PSF=o.img.no_PSF, pt2_data=o.img.data)
    np.savez_compressed(os.path.join(HERE, 'img_analytic.npz'), **out)
    print('img_analytic', out['ser_data'].shape, out['ser_data'].sum(), out['pt_data'].shape, out['pt_data'].sum())


def main():
    ref = ref_loader.load()
    simple = ('simple', {'slope': -1.0})
    sed_case(ref, 'lnu_dust', 2000, 21, simple, simple, 0.0, 7.0, observed=False)
    sed_case(ref, 'fnu_mixed', 1500, 22, ('simple', {'slope': -0.7}), ('Calzetti_like', {}), 0.3, 6.8, observed=True)
    sed_case(ref, 'lnu_nodust', 1000, 23, False, False, 1.0, 7.0, observed=False, with_incident=True)
    sed_case(ref, 'lnu_ismonly_notau', 800, 24, simple, False, 0.1, 7.3, observed=False, pass_tau=False,
             lognormal_mass=False)
    sed_index_case(ref)
    lines_case(ref, 'dust', 1500, 41, simple, ('Calzetti_like', {}), 0.2, 7.0)
    lines_case(ref, 'nodust_incident', 1000, 42, False, False, 0.0, 6.9, with_incident=True)
    lines_case(ref, 'ismonly_notau', 800, 43, simple, False, 0.5, 7.0, pass_tau=False)

    image_case(ref, 'ngp', 3000, 31, 0.1, 50, False)
    image_case(ref, 'convgauss', 3000, 32, 0.1, 50, ('convolved_gaussian', 0.2))
    image_case(ref, 'convgauss_odd', 2000, 36, 0.12, 41, ('convolved_gaussian', 0.35))
    image_case(ref, 'adaptive16', 3000, 33, 0.1, 50, ('adaptive', 16))
    image_case(ref, 'adaptive8f', 2500, 34, 0.07, 64, ('adaptive', 8.))
    image_case(ref, 'adaptive_odd', 1200, 37, 0.09, 77, ('adaptive', 5))

    def few(X, Y, L):  # k > N_sel -> inf distance -> uniform deposit
        return X[:5] * 0.2, Y[:5] * 0.2, L[:5]
    image_case(ref, 'adaptive_kbig', 50, 35, 0.1, 20, ('adaptive', 8), extra=few)

    def coarse(X, Y, L):  # FWHM floor 0.01 with pitch ~1: most particles underflow and are dropped
        X = np.concatenate([X[:40], [0.0, 1.0204081632653061, 0.004, 3.0000001]])
        Y = np.concatenate([Y[:40], [0.0, -1.0204081632653061, 0.003, -2.0]])
        return X * 1.0, Y * 1.0, np.concatenate([L[:40], L[:4]])
    image_case(ref, 'adaptive_underflow', 200, 38, 1.0, 50, ('adaptive', 2), scale_kpc=8.0, extra=coarse)

    pixel_index_case(ref)
    observed_case(ref)
    analytic_case(ref)


if __name__ == '__main__':
    main()
