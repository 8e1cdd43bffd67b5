"""Shared helpers: rebuild the inputs of a golden fixture (tests/golden/*.npz,
written by tests/golden/make_golden.py from the unmodified reference)."""

import ast
import os

import numpy as np

from synthobs_b200 import providers, synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
GRID_KW = dict(n_lam=300, n_age=51, seed=0)
SED_CASES = ['lnu_dust', 'fnu_mixed', 'lnu_nodust', 'lnu_ismonly_notau']
LINES_CASES = ['dust', 'nodust_incident', 'ismonly_notau']
IMG_CASES = ['ngp', 'convgauss', 'convgauss_odd', 'adaptive16', 'adaptive8f', 'adaptive_odd', 'adaptive_kbig',
             'adaptive_underflow']


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False))


def grid_checksum(grid):
    return np.array([grid['stellar'].sum(), grid['nebular'].sum(), grid['log10Q'].sum(),
                     grid['stellar'][7, 3, 100], grid['nebular'][40, 11, 250]])


def golden_grid(g):
    grid = synthetic.ssp_grid(with_incident=bool(g.get('with_incident', False)), **GRID_KW)
    np.testing.assert_allclose(grid_checksum(grid), g['grid_checksum'], rtol=0, atol=0,
                               err_msg='synthetic.ssp_grid no longer reproduces the grid the golden fixtures were made with')
    return grid


def sed_inputs(tag):
    """-> (g, grid, F, dust_ISM, dust_BC, kwargs for the particle functions)"""
    g = load('sed_' + tag)
    grid = golden_grid(g)
    z = float(g['z'])
    if bool(g['observed']):
        F = synthetic.filter_set(grid['lam'] * (1. + z), n_filters=4, kind='gauss', lo=12000., hi=50000.,
                                 prefix='JWST.FAKE.')
    else:
        F = synthetic.filter_set(grid['lam'], n_filters=4, kind='tophat', lo=1500., hi=8000.)
    dust_ISM = ast.literal_eval(str(g['dust_ISM']))
    dust_BC = ast.literal_eval(str(g['dust_BC']))
    kw = dict(fesc=float(g['fesc']), log10t_BC=float(g['log10t_BC']))
    if bool(g['pass_tau']):
        kw.update(tauVs_ISM=g['tauVs_ISM'], tauVs_BC=g['tauVs_BC'])
    return g, grid, F, dust_ISM, dust_BC, kw


def dust_k(dust, lam):
    if not dust:
        return None
    name, params = dust
    return getattr(providers.dust_curves, name)(params=params).tau(lam)


def smoothing_of(g):
    m, s = g['smoothing']
    if str(m) == 'none':
        return False
    s = float(s)
    return (str(m), s)


def lines_inputs(tag):
    """-> (g, lines grid dict, dust_ISM, dust_BC, kwargs) of a lines_* fixture"""
    g = load('lines_' + tag)
    lgrid = synthetic.lines_grid(n_age=GRID_KW['n_age'], seed=0, with_incident=bool(g['with_incident']))
    chk = np.array([lgrid['HI6563']['luminosity'].sum(), lgrid['OII3729']['stellar_continuum'].sum()])
    np.testing.assert_allclose(chk, g['grid_checksum'], rtol=0, atol=0,
                               err_msg='synthetic.lines_grid no longer reproduces the grid of the golden fixtures')
    dust_ISM = ast.literal_eval(str(g['dust_ISM']))
    dust_BC = ast.literal_eval(str(g['dust_BC']))
    kw = dict(fesc=float(g['fesc']), log10t_BC=float(g['log10t_BC']))
    if bool(g['pass_tau']):
        kw.update(tauVs_ISM=g['tauVs_ISM'], tauVs_BC=g['tauVs_BC'])
    return g, lgrid, dust_ISM, dust_BC, kw
