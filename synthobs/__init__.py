"""Drop-in alias: `import synthobs` resolves to the B200 implementation
(synthobs_b200).  Only the hot-path modules exist: synthobs.sed.models,
synthobs.morph.images, synthobs.morph.PSF."""

from synthobs_b200 import empty, __version__  # noqa: F401
from . import sed, morph  # noqa: F401

__all__ = ['morph', 'sed']
