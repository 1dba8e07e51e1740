"""``cna.tl`` — same public names as the reference's ``src/cna/tools/__init__.py:1-10``, plus
``to_device`` (keep the kNN graph resident in HBM between calls)."""
from ._association import association
from ._graph import to_device
from ._nam import diffuse, diffuse_stepwise, nam, svd_nam

__all__ = ["association", "nam", "svd_nam", "diffuse", "diffuse_stepwise", "to_device"]
