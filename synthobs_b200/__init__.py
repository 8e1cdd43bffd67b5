"""synthobs_b200 — B200 (sm_100a) implementation of SynthObs's two data-parallel hot
paths (SED synthesis, particle->image deposition) behind the reference's Python API.

    from synthobs_b200.sed import models
    from synthobs_b200.morph import images

The repo root also carries a `synthobs` alias package so reference scripts that
`import synthobs.sed.models` / `synthobs.morph.images` run unchanged.
"""

__version__ = '0.1.0'


class empty:
    pass
