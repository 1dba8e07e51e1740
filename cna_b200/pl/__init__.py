"""``cna.pl`` mirror — thin, host-only plotting helpers with the reference's names and arguments
(``src/cna/plotting/__init__.py:1-2``: ``umap_ncorr``, ``umap_overlay``, ``violinplot``).  No compute happens
here: they read the per-cell columns ``association()`` wrote to ``data.obs``.  matplotlib (and, for the UMAP
overlays, scanpy and a real AnnData) are imported on first use; neither is needed by anything else in this
package."""
from ._plots import umap_ncorr, umap_overlay, violinplot

__all__ = ["umap_ncorr", "umap_overlay", "violinplot"]
