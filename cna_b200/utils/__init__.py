"""``cna.ut`` mirror: sample-level helpers (reference ``src/cna/utils/__init__.py``)."""
from ._sample_meta import obs_to_sample

__all__ = ["obs_to_sample"]
