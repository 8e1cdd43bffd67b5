"""TEST INFRASTRUCTURE — CPU (NumPy, float64) restatement of the reference's SED
hot path.  Not shipped, not a fallback: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package.

Parity status: the reference's own tests pin nothing for this path (all five
tests/*.py are plot scripts without asserts), so this restatement
is pinned by running the UNMODIFIED reference modules in this container
(oracle/ref_loader.py) on synthetic inputs: tests/golden/make_golden.py wrote
tests/golden/*.npz from the reference, tests/test_oracle_golden.py checks this
file against them.  Third-party arithmetic (interrogator dust curves / IGM,
flare filters) enters only as host-computed inputs: "parity unpinned" there.

Every function cites the reference lines it follows (paths relative to
/root/reference).  Where the reference loops over particles in Python, the
restatement is vectorised over particles in chunks; the per-particle arithmetic
(operation order inside one particle) is kept, the cross-particle float64
summation order is not (difference ~1e-16 relative).
"""

import numpy as np

COMPONENTS = ('stellar_incident', 'stellar_transmitted', 'nebular')


def _trapz(y, x, axis=-1):
    return np.trapezoid(y, x=x, axis=axis)


def ensure_incident(grid):
    """synthobs/sed/models.py:30-35 — derive incident/transmitted from 'stellar'
    when the grid pickle has no 'stellar_incident' key."""
    if 'stellar_incident' not in grid:
        grid['stellar_incident'] = grid['stellar'].copy()
        grid['stellar_transmitted'] = grid['stellar'].copy()
        grid['stellar_transmitted'][:, :, grid['lam'] < 912] = 0.0
    return grid


def ngp_indices(log10age_grid, log10Z_grid, Ages, Metallicities, chunk=65536):
    """synthobs/sed/models.py:108-114 (and :174-180, :212-217, :267-273) —
    literal float64 first-minimum argmin over the age and Z axes.
    Returns (ia, iZ, log10age)."""
    Ages = np.asarray(Ages, dtype=np.float64)
    Metallicities = np.asarray(Metallicities, dtype=np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        log10age = np.log10(Ages) + 6.
        log10Z = np.log10(Metallicities)
    n = Ages.shape[0]
    ia = np.empty(n, dtype=np.int64)
    iZ = np.empty(n, dtype=np.int64)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        with np.errstate(invalid='ignore'):
            ia[s:e] = np.abs(log10age_grid[None, :] - log10age[s:e, None]).argmin(axis=1)
            iZ[s:e] = np.abs(log10Z_grid[None, :] - log10Z[s:e, None]).argmin(axis=1)
    return ia, iZ, log10age


def lnu_grid(grid, F):
    """synthobs/sed/models.py:39-49 — create_Lnu_grid.  Returns
    {f: {component: [na, nZ]}}."""
    out = {}
    for f in F['filters']:
        norm = _trapz(F[f].T, F[f].lam)
        out[f] = {c: _trapz(np.multiply(grid[c], F[f].T), F[f].lam, axis=2) / norm for c in COMPONENTS}
    return out


def fnu_grid(grid, F, z, luminosity_distance_cm, madau):
    """synthobs/sed/models.py:53-67 — create_Fnu_grid (nJy per Msol), same
    operation order as the reference expression."""
    out = {}
    for f in F['filters']:
        igm = madau(F[f].lam, z)
        norm = _trapz(F[f].T, F[f].lam)
        out[f] = {c: 1E23 * 1E9 * _trapz(np.multiply(grid[c] * igm, F[f].T), F[f].lam, axis=2) / norm
                  * (1. + z) / (4. * np.pi * luminosity_distance_cm ** 2) for c in COMPONENTS}
    return out


def broadband_array(tab, log10age_grid, log10Z_grid, Masses, Ages, Metallicities,
                    tauVs_ISM=None, tauVs_BC=None, k_ISM=None, k_BC=None, fesc=0.0, log10t_BC=7.):
    """synthobs/sed/models.py:88-135 (generate_Lnu_array) and :153-200
    (generate_Fnu_array) for ONE filter.  `tab` = {component: [na,nZ]} for that
    filter; k_ISM / k_BC = dust-curve value at the filter pivot wavelength, or
    None when model.dust_ISM / model.dust_BC is False (then T = 1.0, :117-120,
    :126-129).  Missing tauV arrays are zeros (:92-93)."""
    Masses = np.asarray(Masses, dtype=np.float64)
    n = Masses.shape[0]
    tauVs_ISM = np.zeros(n) if tauVs_ISM is None else tauVs_ISM
    tauVs_BC = np.zeros(n) if tauVs_BC is None else tauVs_BC
    ia, iZ, log10age = ngp_indices(log10age_grid, log10Z_grid, Ages, Metallicities)
    T_ISM = np.exp(-(tauVs_ISM * k_ISM)) if k_ISM is not None else np.ones(n)
    T_BC = np.exp(-(tauVs_BC * k_BC)) if k_BC is not None else np.ones(n)
    inc = tab['stellar_incident'][ia, iZ]
    tra = tab['stellar_transmitted'][ia, iZ]
    neb = tab['nebular'][ia, iZ]
    with np.errstate(invalid='ignore'):
        young = log10age < log10t_BC
    l_young = Masses * (T_ISM * fesc * inc + (T_ISM * T_BC) * (1. - fesc) * (tra + neb))
    l_old = Masses * T_ISM * inc
    return np.where(young, l_young, l_old)


def broadband(tabs, filters, *args, **kw):
    """synthobs/sed/models.py:74-85 / :139-150 — generate_Lnu / generate_Fnu:
    {f: sum over particles}.  k_ISM / k_BC are dicts f -> value here."""
    k_ISM = kw.pop('k_ISM', None)
    k_BC = kw.pop('k_BC', None)
    out = {}
    for f in filters:
        out[f] = np.sum(broadband_array(tabs[f], *args,
                                        k_ISM=None if k_ISM is None else k_ISM[f],
                                        k_BC=None if k_BC is None else k_BC[f], **kw))
    return out


def log10Q(grid, Masses, Ages, Metallicities):
    """synthobs/sed/models.py:206-221 — generate_log10Q."""
    ia, iZ, _ = ngp_indices(grid['log10age'], grid['log10Z'], Ages, Metallicities)
    return np.log10(np.sum(np.asarray(Masses) * 10 ** (grid['log10Q'][ia, iZ])))


def sed_spectra(grid, Masses, Ages, Metallicities, tauVs_ISM=None, tauVs_BC=None,
                k_ISM_lam=None, k_BC_lam=None, fesc=0.0, log10t_BC=7., chunk=256):
    """synthobs/sed/models.py:229-314 — generate_SED.  k_*_lam = dust curve over
    the full wavelength grid (:254-262) or None when the model has no such dust
    (T = 1.0, :285-289, :300-304).  Returns dict with the five spectra (:306-310)
    and f_esc, tau, tau_relV (:312-314)."""
    lam = grid['lam']
    Masses = np.asarray(Masses, dtype=np.float64)
    n = Masses.shape[0]
    tauVs_ISM = np.zeros(n) if tauVs_ISM is None else tauVs_ISM
    tauVs_BC = np.zeros(n) if tauVs_BC is None else tauVs_BC
    ia, iZ, log10age = ngp_indices(grid['log10age'], grid['log10Z'], Ages, Metallicities)
    with np.errstate(invalid='ignore'):
        young = log10age < log10t_BC
    out = {k: np.zeros(lam.shape) for k in ('stellar', 'intrinsic', 'BC', 'total', 'no_HII_no_BC')}
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        M = Masses[s:e, None]
        inc = grid['stellar_incident'][ia[s:e], iZ[s:e]]
        tn = grid['stellar_transmitted'][ia[s:e], iZ[s:e]] + grid['nebular'][ia[s:e], iZ[s:e]]
        y = young[s:e, None]
        sed_stellar = M * inc                                                    # :277
        if k_BC_lam is not None:
            T_BC = np.exp(-(tauVs_BC[s:e, None] * k_BC_lam[None, :]))            # :285-287
        else:
            T_BC = 1.0
        intrinsic_y = M * (fesc * inc + (1. - fesc) * tn)                        # :281
        BC_y = M * (fesc * inc + T_BC * (1. - fesc) * tn)                        # :291
        sed_intrinsic = np.where(y, intrinsic_y, M * inc)                        # :295
        sed_BC = np.where(y, BC_y, M * inc)                                      # :296
        if k_ISM_lam is not None:
            T = np.exp(-(tauVs_ISM[s:e, None] * k_ISM_lam[None, :]))             # :300-302
        else:
            T = 1.0
        out['stellar'] += sed_stellar.sum(axis=0)                                # :306
        out['intrinsic'] += sed_intrinsic.sum(axis=0)                            # :307
        out['BC'] += sed_BC.sum(axis=0)                                          # :308
        out['total'] += (T * sed_BC).sum(axis=0)                                 # :309
        out['no_HII_no_BC'] += (T * sed_stellar).sum(axis=0)                     # :310
    with np.errstate(divide='ignore', invalid='ignore'):
        out['f_esc'] = out['total'] / out['intrinsic']                           # :312
        out['tau'] = -np.log(out['f_esc'])                                       # :313
        out['tau_relV'] = out['tau'] / np.interp(5500., lam, out['tau'])         # :314
    return out


def cell_histogram(log10age_grid, log10Z_grid, Masses, Ages, Metallicities, log10t_BC=7.):
    """The grid-weight histogram the north-star formulation feeds to the GEMMs
    (SURVEY.md §8a mapping table): H[2, na, nZ], H[0] = old, H[1] = young,
    H[y, ia, iZ] = sum of Masses of the particles that fall in that cell."""
    na, nZ = log10age_grid.size, log10Z_grid.size
    ia, iZ, log10age = ngp_indices(log10age_grid, log10Z_grid, Ages, Metallicities)
    with np.errstate(invalid='ignore'):
        young = (log10age < log10t_BC).astype(np.int64)
    H = np.zeros((2, na, nZ))
    np.add.at(H, (young, ia, iZ), np.asarray(Masses, dtype=np.float64))
    return H


def line_luminosity(lgrid, line, Masses, Ages, Metallicities, tauVs_ISM=None, tauVs_BC=None,
                    k_ISM=None, k_BC=None, fesc=0.0, log10t_BC=7.):
    """synthobs/sed/models.py:382-448 -- EmissionLines.get_line_luminosity for one (possibly
    blended 'a,b') line.  `lgrid` is the lines.p dict AFTER the constructor's derivation of
    the incident/transmitted continua (:363-367); k_ISM / k_BC = dust-curve value at the mean
    line wavelength (:401,418,426) or None when that dust model is off (T = 1.0)."""
    names = line.split(',')
    Masses = np.asarray(Masses, dtype=np.float64)
    n = Masses.shape[0]
    tauVs_ISM = np.zeros(n) if tauVs_ISM is None else tauVs_ISM                                 # :386-392
    tauVs_BC = np.zeros(n) if tauVs_BC is None else tauVs_BC
    lam = np.mean([lgrid[l]['lam'] for l in names])                                             # :401
    ia, iZ, log10age = ngp_indices(lgrid['log10age'], lgrid['log10Z'], Ages, Metallicities)     # :411-416
    T_ISM = np.exp(-(tauVs_ISM * k_ISM)) if k_ISM is not None else np.ones(n)                   # :418-422
    T_BC = np.exp(-(tauVs_BC * k_BC)) if k_BC is not None else np.ones(n)                       # :426-430
    with np.errstate(invalid='ignore'):
        young = log10age < log10t_BC                                                            # :424
    lum = 0.0
    cont = 0.0
    for l in names:
        g = lgrid[l]
        inc = g['stellar_incident_continuum'][ia, iZ]
        tn = g['stellar_transmitted_continuum'][ia, iZ] + g['nebular_continuum'][ia, iZ]
        lum += np.sum(np.where(young, Masses * T_BC * T_ISM * (1. - fesc) * 10 ** g['luminosity'][ia, iZ], 0.0))   # :433
        cont += np.sum(np.where(young, Masses * (T_ISM * fesc * inc + (T_ISM * T_BC) * (1 - fesc) * tn),          # :434
                                Masses * T_ISM * inc))                                                            # :438
    o = {'luminosity': lum, 'continuum': cont}
    total_continuum = (cont / float(len(names))) * (3E8) / ((lam / float(len(names))) ** 2 * 1E-10)              # :442
    o['EW'] = lum / total_continuum                                                                               # :444
    return o


def ensure_line_continua(lgrid):
    """synthobs/sed/models.py:363-367."""
    if 'stellar_transmitted_continuum' not in lgrid[lgrid['lines'][0]]:
        for l in lgrid['lines']:
            lgrid[l]['stellar_transmitted_continuum'] = lgrid[l]['stellar_continuum']
            lgrid[l]['stellar_incident_continuum'] = lgrid[l]['stellar_continuum']
    return lgrid
