"""Random small configurations: broadband sums / arrays / batch / spectra vs the oracle.  Run on a B200."""
import contextlib, io, sys
import numpy as np
from oracle import sed_oracle
from synthobs_b200 import providers, synthetic
from synthobs_b200.sed import models


def run(seed=0, cases=40):
    """returns the list of failing case numbers (details are printed)"""
    rng = np.random.default_rng(seed)
    bad = []
    for case in range(cases):
        n_lam = int(rng.choice([50, 203, 400]))
        n_age = int(rng.choice([5, 21, 51]))
        grid = synthetic.ssp_grid(n_lam=n_lam, n_age=n_age, seed=int(rng.integers(100)), with_incident=bool(rng.random() < 0.5))
        with contextlib.redirect_stdout(io.StringIO()):
            model = models.define_model(grid)
        dI = ('simple', {'slope': float(rng.choice([-1.0, -0.5, -1.3]))}) if rng.random() < 0.7 else False
        dB = ('Calzetti_like', {}) if rng.random() < 0.5 else (('simple', {'slope': -1.0}) if rng.random() < 0.5 else False)
        model.dust_ISM, model.dust_BC = dI, dB
        nf = int(rng.choice([1, 2, 3, 7, 10, 13, 33, 40]))
        F = synthetic.filter_set(grid['lam'], n_filters=nf, kind=str(rng.choice(['tophat', 'gauss'])), lo=1200., hi=20000.)
        model.create_Lnu_grid(F)
        n = int(rng.choice([0, 1, 2, 31, 1000, 1025, 5000]))
        p = synthetic.particles(max(n, 1), seed=int(rng.integers(1000)), lognormal_mass=bool(rng.random() < 0.5))
        p = {k: v[:n] for k, v in p.items()}
        if n > 3 and rng.random() < 0.3:
            p['Ages'][:3] = [0.0, np.inf, 1e-300]
            p['Metallicities'][:3] = [0.0, 1.0, np.inf]
        fesc = float(rng.choice([0.0, 0.1, 1.0])); t_bc = float(rng.choice([6.8, 7.0, 7.3]))
        kw = dict(fesc=fesc, log10t_BC=t_bc)
        pass_tau = rng.random() < 0.7
        if pass_tau:
            kw.update(tauVs_ISM=p['tauVs_ISM'], tauVs_BC=p['tauVs_BC'])
        og = sed_oracle.ensure_incident({k: (v.copy() if hasattr(v, 'copy') else v) for k, v in grid.items()})
        tabs = sed_oracle.lnu_grid(og, F)
        dk = lambda d, lam: None if not d else getattr(providers.dust_curves, d[0])(params=d[1]).tau(lam)   # noqa: E731
        kI = {f: dk(dI, F[f].pivwv()) for f in F['filters']}; kB = {f: dk(dB, F[f].pivwv()) for f in F['filters']}
        try:
            with np.errstate(all='ignore'):
                ref = sed_oracle.broadband(tabs, F['filters'], grid['log10age'], grid['log10Z'], p['Masses'], p['Ages'],
                                           p['Metallicities'], tauVs_ISM=kw.get('tauVs_ISM'), tauVs_BC=kw.get('tauVs_BC'),
                                           k_ISM=kI if dI else None, k_BC=kB if dB else None, fesc=fesc, log10t_BC=t_bc)
                got = models.generate_Lnu(model, F, p['Masses'], p['Ages'], p['Metallicities'], **kw)
            r = np.array([ref[f] for f in F['filters']]); g = np.array([got[f] for f in F['filters']])
            ok = np.allclose(g, r, rtol=1e-4, atol=1e-4 * np.abs(r).max() * 1e-3 if r.size and np.isfinite(r).all() else 0, equal_nan=True)
            if n > 0:
                f0 = F['filters'][int(rng.integers(nf))]
                with np.errstate(all='ignore'):
                    ra = sed_oracle.broadband_array(tabs[f0], grid['log10age'], grid['log10Z'], p['Masses'], p['Ages'], p['Metallicities'],
                                                    tauVs_ISM=kw.get('tauVs_ISM'), tauVs_BC=kw.get('tauVs_BC'), k_ISM=kI[f0] if dI else None,
                                                    k_BC=kB[f0] if dB else None, fesc=fesc, log10t_BC=t_bc)
                    ga = models.generate_Lnu_array(model, F, f0, p['Masses'], p['Ages'], p['Metallicities'], **kw)
                fin = np.isfinite(ra)
                ok = ok and np.allclose(ga[fin], ra[fin], rtol=2e-5, atol=1e-6 * (np.abs(ra[fin]).max() if fin.any() else 0))
            if not ok:
                bad.append(case)
                print('MISMATCH case', case, dict(n=n, nf=nf, n_age=n_age, dI=dI, dB=dB, fesc=fesc, t_bc=t_bc, tau=pass_tau), np.abs(g / r - 1).max() if r.size else None)
        except Exception as e:   # noqa: BLE001
            bad.append(case)
            print('ERROR case', case, dict(n=n, nf=nf, n_age=n_age), repr(e)[:300])
    return bad



if __name__ == '__main__':
    failures = run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 60)
    print('fuzz done, failures:', len(failures))
