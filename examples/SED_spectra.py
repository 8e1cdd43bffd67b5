"""Counterpart of the reference's examples/SED_spectra.py: full spectra with LOS dust, then band values from them."""
from _common import models, np, quiet, synthetic, synthobs, write_test_object

with quiet():
    model = models.define_model(synthetic.ssp_grid(n_lam=9000))
model.dust_ISM = ('simple', {'slope': -1.0})
model.dust_BC = ('simple', {'slope': -1.0})
test = synthobs.test_data(write_test_object())
test.tauVs_ISM = (10 ** 5.2) * test.MetSurfaceDensities
test.tauVs_BC = 2.0 * (test.Metallicities / 0.01)

with quiet():
    o = models.generate_SED(model, test.Masses, test.Ages, test.Metallicities, tauVs_ISM=test.tauVs_ISM,
                            tauVs_BC=test.tauVs_BC, fesc=0.0)
i1500 = np.argmin(np.abs(o.lam - 1500.))
for name in ('stellar', 'intrinsic', 'BC', 'total', 'no_HII_no_BC'):
    print(f'{name:>13s}: log10 L_nu(1500 A) = {np.log10(getattr(o, name).lnu[i1500]):.3f}')
print('A_1500 =', 2.5 * np.log10(o.intrinsic.lnu[i1500] / o.total.lnu[i1500]), ' tau_relV(1500) =', o.tau_relV[i1500])

F = synthetic.filter_set(model.lam, n_filters=3, kind='tophat', lo=1500., hi=5500.)
print('get_Lnu(F) on the total spectrum:', o.total.get_Lnu(F))                       # host provider (core.sed)
print('same on the device (spectra x filter weights):', dict(zip(F['filters'], models.spectra_Lnu(model, F, o.total.lnu))))
