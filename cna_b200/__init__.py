"""cna_b200 — B200-native covarying neighborhood analysis (NAM + permutation association).

Drop-in for the hot path of ``immunogenomics/cna``::

    import cna_b200 as cna
    p = cna.tl.association(adata, y, 'id', batches=..., covs=...)

``tl`` (``association``, ``nam``, ``svd_nam``, ``diffuse``, ``diffuse_stepwise``) and the host-only
``ut.obs_to_sample`` helper are provided; plotting stays with the reference package.  Importing this module does not load CUDA; the shared
library is loaded (and must exist — there is no CPU fallback) on the first call.
"""
from . import tl  # noqa: F401
from . import utils as ut  # noqa: F401

__version__ = "0.1.0"
