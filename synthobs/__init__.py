"""Drop-in alias: `import synthobs` resolves to the B200 implementation
(synthobs_b200).  Only the hot-path modules exist: synthobs.sed.models,
synthobs.morph.images, synthobs.morph.PSF, synthobs.core (particle-schema loaders)."""

from synthobs_b200 import empty, __version__  # noqa: F401
from . import sed, morph  # noqa: F401
from synthobs_b200 import core  # noqa: F401
from synthobs_b200.core import test_data, get_object_data, all_test_data, Euclid_test_data, FLARE_dir  # noqa: F401  (upstream: from .core import *)
import sys as _sys
_sys.modules[__name__ + '.core'] = core

__all__ = ['core', 'morph', 'sed']
