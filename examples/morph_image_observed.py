"""Counterpart of the reference's examples/morph_image_observed.py: observed-frame images of one object in several
filters (adaptive smoothing, PSF, super-sampling), from per-particle fluxes computed by generate_Fnu_array."""
from _common import models, np, quiet, synthetic, synthobs, write_test_object

import synthobs.morph.images as images

z = 8.0
cosmo = synthetic.FixedCosmology()
with quiet():
    model = models.define_model(synthetic.ssp_grid(n_lam=9000))
model.dust_ISM = ('simple', {'slope': -1.0})
test = synthobs.test_data(write_test_object())
test.tauVs = (10 ** 5.2) * test.MetSurfaceDensities

filters = ['JWST.NIRCAM.F150W', 'JWST.NIRCAM.F200W', 'JWST.NIRCAM.F356W']
for f, ps in zip(filters, (0.031, 0.031, 0.063)):
    images.filter_info[f] = {'pixel_scale': ps}                     # upstream: flare.observatories.filter_info
F = synthetic.filter_set(model.lam * (1. + z), n_filters=3, kind='gauss', lo=15000., hi=35600., prefix='JWST.NIRCAM.F')
F = {'filters': filters, **{f: F[g] for f, g in zip(filters, F['filters'])}}
PSFs = {f: synthetic.GaussianPSF(2.0) for f in filters}
model.create_Fnu_grid(F, z, cosmo)

Fnu = {f: models.generate_Fnu_array(model, F, f, test.Masses, test.Ages, test.Metallicities, tauVs_ISM=test.tauVs,
                                    fesc=0.0) for f in filters}     # nJy per star particle
imgs = images.particle(test.X, test.Y, Fnu, filters, cosmo, z, 3.0, smoothing=('adaptive', 8.), PSFs=PSFs,
                       super_sampling=2)
for f in filters:
    print('{0}: image {1}, total flux {2:.2f} nJy (particles: {3:.2f})'.format(f, imgs[f].data.shape, np.sum(imgs[f].data),
                                                                        np.sum(Fnu[f])))
