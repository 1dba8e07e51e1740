"""Device residency of the kNN graph and sample bookkeeping.

Reference: ``src/cna/tools/_nam.py:12-19`` (get_connectivity), ``:28`` (column sums + self weight),
``:51-54`` (one-hot sample indicator and cells-per-sample counts).
"""
import warnings

import numpy as np
import pandas as pd
import scipy.sparse as sp
import torch

from .. import _lib


def get_connectivity(data):
    """``_nam.py:12-19``: modern AnnData keeps the graph in ``.obsp``, anndata < 0.7.2 in ``.uns``.
    Duck-typed: anything with ``.obsp['connectivities']`` (or the legacy location) works."""
    obsp = getattr(data, "obsp", None)
    if obsp is not None and "connectivities" in obsp:
        return obsp["connectivities"]
    uns = getattr(data, "uns", None)
    if uns is not None and "neighbors" in uns and "connectivities" in uns["neighbors"]:
        return uns["neighbors"]["connectivities"]
    raise KeyError("no kNN graph found: expected data.obsp['connectivities']")


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("cna_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(arr, dtype=None):
    arr = np.ascontiguousarray(arr)
    with warnings.catch_warnings():  # read-only numpy views (pandas CoW, mmap) are only read from
        warnings.simplefilter("ignore", UserWarning)
        t = torch.from_numpy(arr)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(device(), non_blocking=True)


class DeviceGraph:
    """CSR adjacency resident in HBM: int32 indptr / indices plus the raw edge data.  The
    normalised edge values ``A_ij / (colsum_j + w)`` and the diagonal ``w / (colsum_i + w)`` are
    derived per (self_weight, dtype) and cached."""

    def __init__(self, A, shard=None):
        """``shard`` = (comm, row0, row1, rows_per) keeps only rows [row0, row1) on this device
        (cell-axis sharding, ``cna_b200.sharded``); column indices stay global."""
        if not sp.issparse(A):
            raise TypeError("connectivities must be a scipy sparse matrix")
        A = A.tocsr()
        if A.shape[0] != A.shape[1]:
            raise ValueError("connectivities must be square")
        self.n_total = A.shape[0]
        if shard is None:
            self.comm, self.row0, self.rows_per = None, 0, A.shape[0]
            indptr, indices, data = A.indptr, A.indices, A.data
        else:
            from ..sharded import slice_csr
            self.comm, self.row0, row1, self.rows_per = shard
            indptr, indices, data = slice_csr(A, self.row0, row1)
        if len(indices) >= 2 ** 31:
            raise ValueError("graphs with >= 2^31 stored edges per GPU must be sharded further")
        self.n = len(indptr) - 1  # rows held by this device
        self.nnz = int(len(indices))
        self.indptr = _to_dev(indptr, torch.int32)
        self.indices = _to_dev(indices, torch.int32)
        if data.dtype not in (np.float32, np.float64):
            data = data.astype(np.float64)
        self.data = _to_dev(data)
        self._scaled = {}

    def scaled(self, self_weight=1, dtype=torch.float32):
        key = (float(self_weight), dtype)
        if key not in self._scaled:
            colsum = torch.zeros(self.n_total, dtype=torch.float64, device=self.indptr.device)
            _lib.graph_colsum(self.indptr, self.indices, self.data, colsum)
            if self.comm is not None:  # column sums need every shard's rows (_nam.py:28)
                self.comm.all_reduce(colsum)
            vals = torch.empty(self.nnz, dtype=dtype, device=colsum.device)
            diag = torch.empty(self.n, dtype=dtype, device=colsum.device)
            _lib.graph_scale(self.indptr, self.indices, self.data, colsum, self_weight, vals, diag,
                             row_offset=self.row0)
            self._scaled[key] = (vals, diag)
        return self._scaled[key]


class ResidentData:
    """An AnnData-like view whose kNN graph already lives on the GPU (``cna.tl.to_device``).
    ``obs`` is shared with the wrapped object, so ``association`` writes its columns there."""

    def __init__(self, data):
        self._host = data
        self.obsp = getattr(data, "obsp", None)
        self.uns = getattr(data, "uns", None)
        self.graph = DeviceGraph(get_connectivity(data))
        self._codes = {}

    @property
    def obs(self):
        return self._host.obs

    def __len__(self):
        return self.graph.n


def to_device(data):
    """Upload ``data``'s kNN graph once; pass the result to ``nam`` / ``association`` / ``diffuse``
    in place of ``data`` to skip the host-to-device copy on every call."""
    return data if isinstance(data, ResidentData) else ResidentData(data)


def graph_of(data):
    g = getattr(data, "graph", None)
    if isinstance(g, DeviceGraph):  # ResidentData / ShardedData
        return g
    return DeviceGraph(get_connectivity(data))


def sample_codes(data, sid_name):
    """Column order of ``pd.get_dummies(data.obs[sid_name])`` (``_nam.py:51``): the categories of a
    categorical column, otherwise the sorted unique values.  Returns (labels Index, int32 codes
    tensor on the device, cells-per-sample counts as float64 numpy)."""
    cache = getattr(data, "_codes", None)
    sid = data.obs[sid_name]
    categorical = isinstance(sid.dtype, pd.CategoricalDtype)
    raw = sid.cat.codes.to_numpy() if categorical else sid.to_numpy()
    # cache key for resident data: same column buffer => same codes (obs is shared with the host object)
    key = (sid_name, raw.__array_interface__["data"][0], raw.shape[0], str(raw.dtype))
    if cache is not None and key in cache:
        return cache[key]
    if categorical:
        labels = pd.Index(sid.cat.categories)
        codes = raw
        if (codes < 0).any():
            raise ValueError(f"data.obs['{sid_name}'] contains missing values")
    else:
        codes, labels = pd.factorize(sid, sort=True)
        labels = pd.Index(labels)
    counts = np.bincount(codes, minlength=len(labels)).astype(np.float64)
    out = (labels, _to_dev(codes.astype(np.int32)), counts)
    if cache is not None:
        cache.clear()
        cache[key] = out
    return out
