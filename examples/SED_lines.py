"""Counterpart of the reference's examples/SED_lines.py: emission-line luminosities and equivalent widths."""
from _common import models, np, quiet, synthetic, synthobs, write_test_object

with quiet():
    m = models.EmissionLines(synthetic.lines_grid(), dust_BC=('simple', {'slope': -1.0}),
                             dust_ISM=('simple', {'slope': -1.0}), verbose=False)
test = synthobs.test_data(write_test_object())
test.tauVs_ISM = (10 ** 5.2) * test.MetSurfaceDensities
test.tauVs_BC = 2.0 * (test.Metallicities / 0.01)

with quiet():
    o = m.get_line_luminosity('HI6563', test.Masses, test.Ages, test.Metallicities, tauVs_BC=test.tauVs_BC,
                              tauVs_ISM=test.tauVs_ISM)
print('HI6563', {k: float(np.log10(v)) for k, v in o.items()})
with quiet():
    o = m.get_line_luminosities(['HI4861', 'OIII4959,OIII5007', 'OII3726,OII3729'], test.Masses, test.Ages,
                                test.Metallicities, tauVs_BC=test.tauVs_BC, tauVs_ISM=test.tauVs_ISM)
for line, q in o.items():
    print(line, {k: round(float(np.log10(v)), 3) for k, v in q.items()})
