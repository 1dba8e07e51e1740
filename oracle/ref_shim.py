"""TEST INFRASTRUCTURE — loader for the *unmodified* reference package, in-container only.

``/root/reference`` exists only in the build container (never on the GPU box), so this module is
used exclusively by ``tests/golden/make_golden.py`` (fixture generation) and by the CPU tests that
cross-check ``oracle/cna_oracle.py`` against the real reference when it is present.  Nothing in the
product package, ``bench.py`` or the ``-m gpu`` tests imports it.

The reference cannot be imported normally here: ``cna/__init__.py`` pulls in
matplotlib/scanpy, ``_nam.py:5`` imports ``anndata`` and asks for its installed version at
``_nam.py:13``, and ``_association.py:230`` uses ``np.NaN`` (gone in numpy 2).  The shim below stubs
those four things without touching the reference sources.
"""
import importlib
import importlib.metadata as _md
import os
import sys
import types

import numpy as np

REFERENCE_SRC = os.environ.get("CNA_REFERENCE_SRC", "/root/reference/src")


def available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, "cna", "tools"))


def load():
    """Return the reference's ``cna.tools`` module (association, nam, svd_nam, diffuse...)."""
    if not available():
        raise RuntimeError("reference sources not present (expected only in the build container)")
    if "cna.tools" in sys.modules and getattr(sys.modules["cna"], "_b200_shim", False):
        return sys.modules["cna.tools"]
    pkg = types.ModuleType("cna")
    pkg.__path__ = [os.path.join(REFERENCE_SRC, "cna")]
    pkg._b200_shim = True
    sys.modules["cna"] = pkg
    sys.modules.setdefault("anndata", types.ModuleType("anndata"))
    real_version = _md.version

    def _version(name):
        if name == "anndata":
            return "0.10.0"
        return real_version(name)

    _md.version = _version
    if not hasattr(np, "NaN"):
        np.NaN = np.nan
    return importlib.import_module("cna.tools")


class AnnDataLike:
    """The only AnnData surface the hot path touches: ``.obs`` and ``.obsp['connectivities']``."""

    def __init__(self, obs, connectivities):
        self.obs = obs
        self.obsp = {"connectivities": connectivities}
        self.uns = {"neighbors": {"connectivities": connectivities}}

    @property
    def n_obs(self):
        return len(self.obs)
