from synthobs_b200.sed import models  # noqa: F401
import sys as _sys
_sys.modules[__name__ + '.models'] = models
