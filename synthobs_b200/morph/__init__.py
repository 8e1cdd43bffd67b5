from . import images  # noqa: F401
from . import PSF  # noqa: F401
