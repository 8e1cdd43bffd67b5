"""Counterpart of the reference's examples/SED_luminosities.py on a synthetic SSP grid and object: rest-frame
broadband luminosities with the dust terms switched on one after the other."""
from _common import models, np, quiet, synthetic, synthobs, write_test_object

with quiet():
    model = models.define_model(synthetic.ssp_grid(n_lam=9000))       # upstream passes a grid name and path
model.dust_ISM = ('simple', {'slope': -1.0})
model.dust_BC = ('simple', {'slope': -1.0})

test = synthobs.test_data(write_test_object())
tau_BC = 2.0 * (test.Metallicities / 0.01)
tau_ISM = (10 ** 5.2) * test.MetSurfaceDensities

F = synthetic.filter_set(model.lam, n_filters=3, kind='tophat', lo=1500., hi=5500.)   # stands in for flare.filters
model.create_Lnu_grid(F)                                              # erg/s/Hz per Msol: one tensor-core contraction

cases = [('Pure stellar luminosities', dict(fesc=1.0)),
         ('With just HII region', dict(fesc=0.0)),
         ('With just BC dust', dict(fesc=0.0, tauVs_BC=tau_BC)),
         ('Total', dict(fesc=0.0, tauVs_BC=tau_BC, tauVs_ISM=tau_ISM))]
for title, kw in cases:
    Lnu = models.generate_Lnu(model, F, test.Masses, test.Ages, test.Metallicities, **kw)
    print('--- ' + title)
    print('   ', '  '.join('%s %.2f' % (f, np.log10(Lnu[f])) for f in F['filters']))
print('log10Q', models.generate_log10Q(model, test.Masses, test.Ages, test.Metallicities))
