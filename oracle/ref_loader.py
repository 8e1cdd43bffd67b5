"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference (/root/reference/synthobs)
in-process so its own functions can be run on synthetic inputs.

Used only by tests/golden/make_golden.py and the `ref`-marked CPU tests, in the
build container; /root/reference does not exist on the GPU box and nothing in
the product, the -m gpu tests, smoke() or bench.py imports this module.

`import synthobs` fails as-is (synthobs/core.py:6 `import flare`), so stub
modules are injected into sys.modules for the packages that are not installed
here: flare, matplotlib, astropy, interrogator.  The numeric
stubs (dust_curves, IGM.madau, core.sed, convolve_fft) forward to
synthobs_b200.providers / the restated astropy identity below, i.e. the SAME
host-side numbers the CUDA path is fed, so parity does not depend on the stubs'
fidelity to upstream.  One shim is unavoidable: scipy>=1.6 renamed
cKDTree.query(n_jobs=) to workers= (synthobs/morph/images.py:85), and
k may arrive as a float (tests/test_images_core.py:28).

The reference package is loaded under the name `synthobs_reference` so it cannot
collide with this repo's drop-in `synthobs` alias package.
"""

import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get('SYNTHOBS_REFERENCE', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'synthobs', '__init__.py'))


def convolve_fft(array, kernel, **kw):
    """astropy.convolution.convolve_fft with its defaults, restated [recall,
    unverifiable offline — parity unpinned at this third-party boundary]:
    boundary='fill' (zeros), normalize_kernel=True, kernel centred on element
    n//2 (bigkernel slice centre - n//2 .. then ifftshift), output cropped to
    the array's own slice.  Identity: out[j,i] = sum_ab img[a,b] *
    khat[j-a+ky//2, i-b+kx//2], khat = k/sum(k), zero outside."""
    import scipy.fft as sfft
    a = np.asarray(array, dtype=np.float64)
    k = np.asarray(kernel, dtype=np.float64)
    k = k / k.sum()
    shape = [sfft.next_fast_len(a.shape[d] + k.shape[d] - 1, real=True) for d in range(2)]
    full = sfft.irfft2(sfft.rfft2(a, shape) * sfft.rfft2(k, shape), shape)
    oy, ox = k.shape[0] // 2, k.shape[1] // 2
    return full[oy:oy + a.shape[0], ox:ox + a.shape[1]].copy()


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


# filter_info used by observed.__init__ (images.py:125); pixel scales in arcsec
FILTER_INFO = {
    'JWST.NIRCAM.F115W': {'pixel_scale': 0.031}, 'JWST.NIRCAM.F150W': {'pixel_scale': 0.031},
    'JWST.NIRCAM.F200W': {'pixel_scale': 0.031}, 'JWST.NIRCAM.F277W': {'pixel_scale': 0.063},
    'JWST.NIRCAM.F356W': {'pixel_scale': 0.063}, 'JWST.NIRCAM.F444W': {'pixel_scale': 0.063},
    'HST.WFC3.F160W': {'pixel_scale': 0.13}, 'HST.WFC3.F125W': {'pixel_scale': 0.13},
}


def _install_stubs():
    from synthobs_b200 import providers
    import scipy.spatial

    flare = _mod('flare', FLARE_dir='/nonexistent_FLARE')
    flare.filters = _mod('flare.filters', pixel_scale={k: v['pixel_scale'] for k, v in FILTER_INFO.items()})
    flare.observatories = _mod('flare.observatories', filter_info=FILTER_INFO)

    mpl = _mod('matplotlib')
    mpl.pyplot = _mod('matplotlib.pyplot')

    astropy = _mod('astropy')
    astropy.convolution = _mod('astropy.convolution', convolve=None, convolve_fft=convolve_fft,
                               Gaussian2DKernel=None)
    astropy.modeling = _mod('astropy.modeling')
    astropy.modeling.models = _mod('astropy.modeling.models', Sersic2D=providers.Sersic2D, Gaussian2D=None)
    astropy.io = _mod('astropy.io', fits=None)

    interrogator = _mod('interrogator')
    interrogator.sed = _mod('interrogator.sed')
    interrogator.sed.core = _mod('interrogator.sed.core', sed=providers.sed)
    interrogator.sed.IGM = _mod('interrogator.sed.IGM', madau=providers.madau)
    interrogator.sed.dust_curves = _mod('interrogator.sed.dust_curves', simple=providers.simple,
                                        Calzetti_like=providers.Calzetti_like)

    if not getattr(scipy.spatial.cKDTree, '_synthobs_shim', False):
        base = scipy.spatial.cKDTree

        class cKDTree(base):
            _synthobs_shim = True

            def query(self, x, k=1, n_jobs=None, **kw):
                if n_jobs is not None:
                    kw['workers'] = n_jobs
                return super().query(x, k=int(k), **kw)

        scipy.spatial.cKDTree = cKDTree

    if not hasattr(np, 'trapz'):
        np.trapz = np.trapezoid


_loaded = None


def load():
    """Return the reference package (attributes .sed.models, .morph.images)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('reference not present at %s' % REFERENCE_ROOT)
    _install_stubs()
    pkg_dir = os.path.join(REFERENCE_ROOT, 'synthobs')
    spec = importlib.util.spec_from_file_location(
        'synthobs_reference', os.path.join(pkg_dir, '__init__.py'), submodule_search_locations=[pkg_dir])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules['synthobs_reference'] = pkg
    spec.loader.exec_module(pkg)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        models = importlib.import_module('synthobs_reference.sed.models')
        images = importlib.import_module('synthobs_reference.morph.images')
    pkg.sed.models = models
    pkg.morph.images = images
    _loaded = pkg
    return pkg
