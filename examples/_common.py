import contextlib
import io
import os
import sys
import tempfile

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))

import numpy as np                     # noqa: E402

import synthobs                        # noqa: E402  (the drop-in alias package of this repo)
from synthobs.sed import models        # noqa: E402
from synthobs_b200 import synthetic    # noqa: E402


def write_test_object(n=20000, seed=1):
    """A synthetic object in the on-disk schema synthobs.test_data() reads (synthobs/core.py:13-24)."""
    p = synthetic.particles(n, seed=seed)
    d = tempfile.mkdtemp()
    synthobs.core.write_object(d, p['X'], p['Y'], p['MetSurfaceDensities'], p['Ages'], p['Metallicities'])
    return d


def quiet():
    return contextlib.redirect_stdout(io.StringIO())
