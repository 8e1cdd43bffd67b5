from synthobs_b200.morph import images, PSF  # noqa: F401
import sys as _sys
_sys.modules[__name__ + '.images'] = images
_sys.modules[__name__ + '.PSF'] = PSF
