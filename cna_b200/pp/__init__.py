"""``cna_b200.pp`` — the step upstream of the hot path: the kNN graph CNA diffuses over.

The reference takes ``data.obsp['connectivities']`` from ``scanpy.pp.neighbors`` (``demo/demo.ipynb`` cell 29,
``demo/makedata.ipynb`` cell 5; read at ``src/cna/tools/_nam.py:12-19``).  ``neighbors`` builds the same
object on the GPU: an exact kNN search and UMAP's fuzzy-simplicial-set weights."""
from ._neighbors import fuzzy_simplicial_set, knn, neighbors

__all__ = ["neighbors", "knn", "fuzzy_simplicial_set"]
