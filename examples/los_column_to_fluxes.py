"""No upstream equivalent: the step BEFORE the reference's inputs.  The reference loads MetSurfaceDensities.npy
(synthobs/core.py:19) and turns it into the ISM optical depth (examples/SED_luminosities.py:69); here the column density is
computed from star and gas particles (synthobs_b200.los: gas in front of every star, SPH kernels projected along the line
of sight) and fed into the same broadband call."""
from _common import models, np, quiet, synthetic

from synthobs_b200 import los

rng = np.random.default_rng(4)
n_star, n_gas = 20000, 40000
p = synthetic.particles(n_star, seed=3)                               # X, Y [kpc], Ages [Myr], Metallicities, Masses
z_star = rng.normal(0.0, 1.0, n_star)
gas_xyz = rng.normal(0.0, 2.0, (3, n_gas))
gas_mass = np.full(n_gas, 3.0e4)                                      # M_sol
gas_Z = 10 ** rng.uniform(-3.5, -1.8, n_gas)
gas_h = 10 ** rng.uniform(-1.2, -0.2, n_gas)                          # kernel support radii [kpc]

# scale 1e-10: units of 1e10 M_sol / kpc^2, in which the reference's 10^5.2 gives optical depths of order one
Sigma = los.metal_surface_density(p['X'], p['Y'], z_star, *gas_xyz, gas_mass, gas_Z, gas_h, scale=1e-10)
tau_ISM = (10 ** 5.2) * Sigma                                         # examples/SED_luminosities.py:69
tau_BC = 2.0 * (p['Metallicities'] / 0.01)
print('metal column density [1e10 M_sol / kpc^2]: median %.3e, 90th percentile %.3e, %d of %d stars behind no gas' % (
    np.median(Sigma), np.percentile(Sigma, 90), int((Sigma == 0).sum()), n_star))
print('tau_V (ISM): median %.2f' % np.median(tau_ISM))

with quiet():
    model = models.define_model(synthetic.ssp_grid(n_lam=9000))
model.dust_ISM = ('simple', {'slope': -1.0})
model.dust_BC = ('simple', {'slope': -1.0})
F = synthetic.filter_set(model.lam, n_filters=3, kind='tophat', lo=1500., hi=5500.)
model.create_Lnu_grid(F)
for title, kw in (('no ISM dust', dict(fesc=0.0, tauVs_BC=tau_BC)), ('with the computed ISM dust', dict(fesc=0.0, tauVs_BC=tau_BC, tauVs_ISM=tau_ISM))):
    Lnu = models.generate_Lnu(model, F, p['Masses'], p['Ages'], p['Metallicities'], **kw)
    print('--- ' + title)
    print('   ', '  '.join('%s %.3f' % (f, np.log10(Lnu[f])) for f in F['filters']))
