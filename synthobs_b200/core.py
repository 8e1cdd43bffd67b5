"""synthobs.core -- the on-disk particle schema of the reference loaders
(reference: synthobs/core.py:13-131): one directory per object with X.npy, Y.npy,
MetSurfaceDensities.npy, Ages.npy [Myr], Metallicities.npy (float64), masses fixed at the
BlueTides particle mass.  Same names and attributes as upstream; additionally
`csr_batch` concatenates objects into the CSR layout the batched entry points take
(models.generate_Fnu_batch, generate_SED_batch, ...).  Host-side I/O only.
"""

import os

import numpy as np

try:
    import flare  # type: ignore
    FLARE_dir = flare.FLARE_dir
except Exception:
    FLARE_dir = os.environ.get('FLARE', '')

# mass of each star particle in M_sol, BLUETIDES, h-less units (core.py:22)
PARTICLE_MASS = 1E10 * 5.90556119E-05 / 0.697

FIELDS = ('X', 'Y', 'MetSurfaceDensities', 'Ages', 'Metallicities')


class empty:
    pass


def _load_into(obj, path):
    for f in FIELDS:
        setattr(obj, f, np.load(os.path.join(path, f + '.npy')))
    obj.Masses = np.ones(obj.Ages.shape) * PARTICLE_MASS
    return obj


def test_data(path_to_example_data=FLARE_dir + '/data/package_example_data/SynthObs/test'):
    """reference: core.py:13-24"""
    return _load_into(empty(), path_to_example_data)


class get_object_data():
    """reference: core.py:27-38"""

    def __init__(self, snap, ID, subID, data_path=FLARE_dir + 'simulations/BlueTides/all'):
        _load_into(self, data_path + '/{snap}/{ID}/{subID}/'.format(snap=snap, ID=ID, subID=subID))


class all_test_data():
    """reference: core.py:44-89 -- every object directory <root>/<snap>/<ID>/<subID>, sorted by mass."""

    def __init__(self, snap, path_to_example_data=FLARE_dir + '/data/package_example_data/SynthObs/'):
        self.data_path = path_to_example_data
        self.obj = []
        for x in os.walk(path_to_example_data + snap):
            if x[0].split('/')[-3] == snap:
                self.obj.append('/'.join(x[0].split('/')[-3:]))
        self.M = np.array([len(np.load(self.data_path + o + '/Ages.npy')) * PARTICLE_MASS for o in self.obj])
        order = np.argsort(self.M)
        self.M = self.M[order]
        self.obj = np.array(self.obj)[order]
        self.Mr = np.round(np.log10(self.M), 1)
        self.N = len(self.obj)
        self.j = 0

    def get(self, i):
        data = _load_into(empty(), self.data_path + self.obj[i] + '/')
        data.id = self.obj[i]
        return data

    def next(self):
        self.j += 1
        return self.get(self.j - 1)


class Euclid_test_data(all_test_data):
    """reference: core.py:90-131 -- same walk and loader as all_test_data, default root EuclidPredictions"""

    def __init__(self, snap, path_to_example_data=FLARE_dir + '/data/project_data/EuclidPredictions/'):
        all_test_data.__init__(self, snap, path_to_example_data)


def write_object(path, X, Y, MetSurfaceDensities, Ages, Metallicities):
    """Write one object directory in the schema above (for offline / synthetic catalogues)."""
    os.makedirs(path, exist_ok=True)
    for f, a in zip(FIELDS, (X, Y, MetSurfaceDensities, Ages, Metallicities)):
        np.save(os.path.join(path, f + '.npy'), np.asarray(a, dtype=np.float64))


def csr_batch(objects, tauV_ISM_norm=10 ** 5.2, tauV_BC_norm=2.0):
    """Concatenate loaded objects into one catalogue + CSR `offsets` [G+1] for the batched entry
    points.  Optical depths as in examples/SED_luminosities.py:69-70:
    tauVs_ISM = 10^5.2 * MetSurfaceDensities, tauVs_BC = 2 * Metallicities / 0.01."""
    objects = list(objects)
    sizes = np.array([len(o.Ages) for o in objects], dtype=np.int64)
    offsets = np.zeros(len(objects) + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    out = {f: np.concatenate([getattr(o, f) for o in objects]) if objects else np.zeros(0) for f in FIELDS + ('Masses',)}
    out['offsets'] = offsets
    out['tauVs_ISM'] = tauV_ISM_norm * out['MetSurfaceDensities']
    out['tauVs_BC'] = tauV_BC_norm * (out['Metallicities'] / 0.01)
    return out
