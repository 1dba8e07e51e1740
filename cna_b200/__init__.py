"""cna_b200 — B200-native covarying neighborhood analysis (NAM + permutation association).

Drop-in for the hot path of ``immunogenomics/cna``::

    import cna_b200 as cna
    p = cna.tl.association(adata, y, 'id', batches=..., covs=...)

``tl`` (``association``, ``nam``, ``svd_nam``, ``diffuse``, ``diffuse_stepwise``) and the host-only
``ut.obs_to_sample`` helper are provided, plus the two steps either side of the hot path: ``pp.neighbors``
(the kNN graph ``scanpy.pp.neighbors`` would build, exact search on the GPU) and ``read_h5ad`` /
``read_connectivities`` (``.h5ad`` -> page-locked CSR, whole or one rank's block of rows), and ``pl``, thin
mirrors of the reference's three plotting helpers (matplotlib / scanpy imported on use).  Importing this module does not load CUDA; the shared library is loaded (and
must exist — there is no CPU fallback) on the first call.
"""
from . import pl  # noqa: F401
from . import pp  # noqa: F401
from . import tl  # noqa: F401
from . import utils as ut  # noqa: F401
from .io import read_connectivities, read_h5ad  # noqa: F401

__version__ = "0.1.0"
