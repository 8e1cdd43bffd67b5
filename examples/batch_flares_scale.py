"""A FLARES-like batch (BASELINE config 4) in single calls: broadband fluxes, spectra and images of many galaxies."""
from _common import models, np, quiet, synthetic

import synthobs.morph.images as images

z = 8.0
with quiet():
    model = models.define_model(synthetic.ssp_grid(n_lam=9000))
model.dust_ISM = model.dust_BC = ('simple', {'slope': -1.0})
F = synthetic.filter_set(model.lam * (1. + z), n_filters=10, kind='gauss', lo=12000., hi=50000., prefix='JWST.FAKE.')
model.create_Fnu_grid(F, z, synthetic.FixedCosmology())

b = synthetic.galaxy_batch(200, n_min=1000, n_max=100000, seed=5)       # concatenated particles + CSR offsets
args = (b['Masses'], b['Ages'], b['Metallicities'])
kw = dict(tauVs_ISM=b['tauVs_ISM'], tauVs_BC=b['tauVs_BC'])
fnu = models.generate_Fnu_batch(model, F, *args, b['offsets'], **kw)                 # [G, F] nJy
print('fluxes', fnu.shape, 'median log10 F_nu/nJy per filter:', np.round(np.median(np.log10(fnu), axis=0), 2))
sed = models.generate_SED_batch(model, *args, b['offsets'], **kw)                    # five [G, n_lam] spectra
print('spectra', {k: v.shape for k, v in sed.items()})
L = models.generate_Fnu_arrays(model, F, *args, **kw)[F['filters'][3]]               # per-particle fluxes in one filter
img = images.core_batch(b['X'], b['Y'], L, b['offsets'], resolution=0.1, ndim=100, smoothing=('adaptive', 16))
print('images', img.data.shape, 'flux inside the frames / total:', float(img.data.sum() / L.sum()))
