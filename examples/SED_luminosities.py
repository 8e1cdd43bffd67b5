"""Counterpart of the reference's examples/SED_luminosities.py on a synthetic SSP grid and object."""
from _common import models, np, quiet, synthetic, synthobs, write_test_object

with quiet():
    model = models.define_model(synthetic.ssp_grid(n_lam=9000))       # upstream: define_model('<grid name>', path)
model.dust_ISM = ('simple', {'slope': -1.0})
model.dust_BC = ('simple', {'slope': -1.0})

test = synthobs.test_data(write_test_object())                        # --- read in some test data

F = synthetic.filter_set(model.lam, n_filters=3, kind='tophat', lo=1500., hi=5500.)   # upstream: flare.filters.add_filters
model.create_Lnu_grid(F)                                              # erg/s/Hz per Msol, one tensor-core contraction

print('--- Pure stellar luminosities')
Lnu = models.generate_Lnu(model, F, test.Masses, test.Ages, test.Metallicities, fesc=1.0)
for f in F['filters']:
    print(f'{f} {np.log10(Lnu[f]):.2f}')

print('--- With just HII region')
Lnu = models.generate_Lnu(model, F, test.Masses, test.Ages, test.Metallicities, fesc=0.0)
for f in F['filters']:
    print(f'{f} {np.log10(Lnu[f]):.2f}')

print('--- With just BC dust')
test.tauVs_BC = 2.0 * (test.Metallicities / 0.01)
Lnu = models.generate_Lnu(model, F, test.Masses, test.Ages, test.Metallicities, tauVs_BC=test.tauVs_BC, fesc=0.0)
for f in F['filters']:
    print(f'{f} {np.log10(Lnu[f]):.2f}')

print('--- Total')
test.tauVs_ISM = (10 ** 5.2) * test.MetSurfaceDensities
Lnu = models.generate_Lnu(model, F, test.Masses, test.Ages, test.Metallicities, tauVs_ISM=test.tauVs_ISM,
                          tauVs_BC=test.tauVs_BC, fesc=0.0)
for f in F['filters']:
    print(f'{f} {np.log10(Lnu[f]):.2f}')
print('log10Q', models.generate_log10Q(model, test.Masses, test.Ages, test.Metallicities))
