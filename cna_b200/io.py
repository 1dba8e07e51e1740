"""Ingestion of ``.h5ad`` files for the hot path: the kNN graph and the numeric
``obs`` columns the reference reads (``src/cna/tools/_nam.py:12-19``: ``data.obsp['connectivities']``, or the
legacy ``data.uns['neighbors']['connectivities']``; ``_nam.py:51``: ``data.obs[sid_name]``), without
anndata / h5py (neither is in this image): ``cna_b200.utils.h5min`` maps the file and decodes the subset of
HDF5 that anndata writes for uncompressed files.

    d = cna_b200.read_h5ad("data.h5ad")                        # whole graph, CSR buffers page-locked
    d = cna_b200.read_h5ad("data.h5ad", rows=(r0, r1))         # one rank's block of rows (r1 - r0) x N
    p = cna_b200.tl.association(cna_b200.sharded.shard_to_device(d), y, "id", ...)

The CSR triplet is copied from the mapped file straight into page-locked host memory (one pass), so the
upload that follows runs at PCIe speed, and a row block never touches the rest of the file.
"""
import numpy as np
import scipy.sparse as sp

from .utils.h5min import H5File

_GRAPH_GROUPS = ("obsp/connectivities", "uns/neighbors/connectivities")


def _pinned_array(n, dtype, pin):
    """A numpy array of n elements, page-locked when CUDA is available and ``pin``."""
    if pin:
        try:
            import torch
            if torch.cuda.is_available():
                t = torch.empty(int(n), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
                arr = t.numpy()
                return arr, t  # the tensor owns the page-locked allocation
        except Exception:  # noqa: BLE001 - no CUDA runtime: ordinary memory
            pass
    return np.empty(int(n), dtype=dtype), None


def read_connectivities(path, rows=None, pin=True, file=None):
    """The kNN graph of an ``.h5ad`` file as a scipy CSR (float64 data, int32 indices): all rows, or the block
    ``rows`` = (r0, r1) as an (r1 - r0) x N matrix.  ``pin``: CSR buffers in page-locked memory (kept alive by
    the returned matrix)."""
    f = file or H5File(path)
    members = f.tree()
    group = next((g for g in _GRAPH_GROUPS if g + "/indptr" in members), None)
    if group is None:
        raise KeyError("no kNN graph in the file (looked for " + ", ".join(_GRAPH_GROUPS) + "); "
                       "run cna_b200.pp.neighbors / scanpy.pp.neighbors first")
    indptr = f.read(group + "/indptr").astype(np.int64)
    n_rows = len(indptr) - 1
    r0, r1 = (0, n_rows) if rows is None else (int(rows[0]), int(rows[1]))
    if not 0 <= r0 <= r1 <= n_rows:
        raise ValueError(f"rows={rows} outside [0, {n_rows}]")
    e0, e1 = int(indptr[r0]), int(indptr[r1])
    keep = []
    data, t = _pinned_array(e1 - e0, np.float64, pin)
    keep.append(t)
    shape, dtype, _ = f.dataset_info(group + "/data")
    if dtype == np.float64:
        f.read_slice(group + "/data", e0, e1, out=data)
    else:  # scanpy writes float32 weights: widened here, the reference's arithmetic is float64
        data[...] = f.read_slice(group + "/data", e0, e1)
    indices, t = _pinned_array(e1 - e0, np.int32, pin)
    keep.append(t)
    indices[...] = f.read_slice(group + "/indices", e0, e1)
    ptr, t = _pinned_array(r1 - r0 + 1, np.int32, pin)
    keep.append(t)
    ptr[...] = indptr[r0:r1 + 1] - e0
    A = sp.csr_matrix((data, indices, ptr), shape=(r1 - r0, n_rows), copy=False)
    A._cna_pinned = keep  # the page-locked allocations live as long as the matrix
    return A


def read_h5ad(path, rows=None, obs_columns=None, pin=True):
    """An AnnData-like object (``cna_b200.synth.AnnDataLike``) with the kNN graph in ``.obsp['connectivities']``
    and the numeric / categorical-code columns of ``obs`` (``obs_columns``: names, default all that decode).
    ``rows`` = (r0, r1) reads one block of rows of the graph (``.row_block``); ``obs`` always covers all
    cells (the sample ids of the halo cells are needed by the first diffusion step)."""
    import pandas as pd

    from .synth import AnnDataLike
    f = H5File(path)
    members = f.tree()
    A = read_connectivities(path, rows=rows, pin=pin, file=f)
    n = A.shape[1]
    cols = {}
    names = sorted({k.split("/")[1] for k in members if k.startswith("obs/")})
    for name in names:
        if obs_columns is not None and name not in obs_columns:
            continue
        try:
            if f"obs/{name}/codes" in members:  # anndata >= 0.7 categorical: codes + categories
                codes = f.read(f"obs/{name}/codes")
                cats = f.read(f"obs/{name}/categories")
                if cats.dtype.kind == "S":
                    cats = cats.astype(str)
                cols[name] = pd.Categorical.from_codes(codes, categories=cats)
            elif f"obs/{name}" in members:
                arr = f.read(f"obs/{name}")
                if arr.ndim == 1 and len(arr) == n:
                    cols[name] = arr.astype(str) if arr.dtype.kind == "S" else arr
        except NotImplementedError:
            if obs_columns is not None:
                raise
    if obs_columns is not None:
        missing = [c for c in obs_columns if c not in cols]
        if missing:
            raise KeyError(f"obs columns {missing} are not in the file (or use an HDF5 feature h5min does not decode)")
    data = AnnDataLike(pd.DataFrame(cols, index=pd.RangeIndex(n)), A)
    if rows is not None:
        data.row_block = (int(rows[0]), int(rows[1]))
    return data
