"""Device residency of the kNN graph and sample bookkeeping.

Reference: ``src/cna/tools/_nam.py:12-19`` (get_connectivity), ``:28`` (column sums + self weight),
``:51-54`` (one-hot sample indicator and cells-per-sample counts).
"""
import os
import warnings

import numpy as np
import pandas as pd
import scipy.sparse as sp
import torch

from .. import _lib
from ._timing import mark


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


_HAVE_CUDA = None


def device():
    global _HAVE_CUDA
    if _HAVE_CUDA is None:
        _HAVE_CUDA = bool(torch.cuda.is_available())
        if _HAVE_CUDA:
            torch.cuda.init()
    if not _HAVE_CUDA:
        raise RuntimeError("cna_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    return torch.device("cuda", torch._C._cuda_getDevice())


def _to_dev(arr, dtype=None):
    arr = np.ascontiguousarray(arr)
    if arr.flags.writeable:
        t = torch.from_numpy(arr)
    else:
        with warnings.catch_warnings():  # read-only numpy views (pandas CoW, mmap) are only read from
            warnings.simplefilter("ignore", UserWarning)
            t = torch.from_numpy(arr)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(device(), non_blocking=True)


REORDER_MIN_CELLS = 150_000  # below this the whole diffusion state lives in L2 anyway


def want_reorder(n_cells):
    """Reorder by default only when the state cannot stay in L2; ``CNA_B200_REORDER=0/1`` forces."""
    env = os.environ.get("CNA_B200_REORDER")
    if env is not None:
        return env.strip().lower() not in ("0", "", "false", "no")
    return n_cells >= REORDER_MIN_CELLS


def _bfs_far_node(indptr, indices, n, root, max_levels):
    """Last node discovered by an (unsorted) breadth-first sweep from ``root``: one end of a long
    shortest path, the usual pseudo-peripheral starting point for Cuthill-McKee."""
    dev = indptr.device
    level = torch.full((n,), -1, dtype=torch.int32, device=dev)
    first_parent = torch.full((n,), 2 ** 31 - 1, dtype=torch.int32, device=dev)
    bufs = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(2)]
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    level[root] = 0
    frontier = torch.tensor([root], dtype=torch.int32, device=dev)
    for lvl in range(max_levels):
        count.zero_()
        nxt = bufs[lvl & 1]
        _lib.bfs_expand(indptr, indices, frontier, 0, lvl + 1, level, first_parent, nxt, count)
        c = int(count.item())
        if c == 0:
            break
        frontier = nxt[:c]
    return int(frontier.min().item())  # min: independent of the order of discovery


def cuthill_mckee_order(indptr, indices, n, max_levels=4096, max_roots=8, far_root=True):
    """Cuthill-McKee ordering of a symmetric CSR on the device (csrc/reorder.cu): breadth-first levels
    from a pseudo-peripheral root, each level sorted by (position of its first parent, node id).
    ``far_root=False`` skips the extra sweep that looks for the pseudo-peripheral root (one-shot calls
    on a host graph, where the ordering is on the critical path: a slightly wider band, ~30 % cheaper).
    Returns (order int64 [n] new -> old, inv int32 [n] old -> new), or None when the graph is too
    path-like to be worth it (more than ``max_levels`` levels).  Components beyond ``max_roots`` are
    appended in their original order.  Deterministic: sets and sort keys do not depend on timing."""
    dev = indptr.device
    level = torch.full((n,), -1, dtype=torch.int32, device=dev)
    first_parent = torch.full((n,), 2 ** 31 - 1, dtype=torch.int32, device=dev)
    order = torch.empty(n, dtype=torch.int64, device=dev)
    nxt = torch.empty(n, dtype=torch.int32, device=dev)
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    front_buf = torch.empty(n, dtype=torch.int32, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    deg = (indptr[1:] - indptr[:-1]).long()
    placed, lvl = 0, 0
    for attempt in range(max_roots):
        if placed >= n:
            break
        big = torch.iinfo(torch.int64).max
        root = int(torch.where(level < 0, deg, torch.full_like(deg, big)).argmin().item())
        if attempt == 0 and far_root:  # the giant component: start from the far end of a long shortest path
            root = _bfs_far_node(indptr, indices, n, root, max_levels)
            mark("graph: pseudo-peripheral root found")
        level[root] = lvl
        frontier = torch.tensor([root], dtype=torch.int32, device=dev)
        order[placed] = root
        pos_base, placed = placed, placed + 1
        while True:
            count.zero_()
            _lib.bfs_expand(indptr, indices, frontier, pos_base, lvl + 1, level, first_parent, nxt, count)
            c = int(count.item())
            if c == 0:
                break
            _lib.bfs_keys(nxt, c, first_parent, keys)
            srt = torch.sort(keys[:c]).values
            _lib.bfs_place(srt, c, order, placed, front_buf)
            frontier = front_buf[:c]
            pos_base, placed, lvl = placed, placed + c, lvl + 1
            if lvl > max_levels:
                return None
        lvl += 1
    if placed < n:
        order[placed:] = torch.nonzero(level < 0).reshape(-1)
    inv = torch.empty(n, dtype=torch.int32, device=dev)
    inv[order] = torch.arange(n, dtype=torch.int32, device=dev)
    return order, inv


_UPLOAD_STREAMS = {}


def _upload_stream(dev):
    """One side stream per device for the asynchronous part of graph uploads (a fresh stream per
    call would make the caching allocator keep a separate 0.3 GB block for each of them)."""
    key = (dev.type, dev.index)
    if key not in _UPLOAD_STREAMS:
        _UPLOAD_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _UPLOAD_STREAMS[key]


class DeviceGraph:
    """CSR adjacency resident in HBM: int32 indptr / indices plus the raw edge data.  The
    normalised edge values ``A_ij / (colsum_j + w)`` and the diagonal ``w / (colsum_i + w)`` are
    derived per (self_weight, dtype) and cached.

    Large graphs are stored in a Cuthill-McKee cell order (``order``: new -> old, ``inv``: old ->
    new; both None when the original order is kept).  Everything row-indexed on the device is in the
    new order; the host boundary (``unpermute``) restores the caller's order."""

    def __init__(self, A, shard=None, reorder=None, resident=True):
        """``shard`` = (comm, row0, row1, rows_per) keeps only rows [row0, row1) (of the stored
        order) on this device (cell-axis sharding, ``cna_b200.sharded``); column indices stay
        global."""
        if not sp.issparse(A):
            raise TypeError("connectivities must be a scipy sparse matrix")
        A = A.tocsr()
        block_only = shard is not None and shard[0].world > 1 and A.shape[0] != A.shape[1]
        if A.shape[0] != A.shape[1] and not block_only:
            raise ValueError("connectivities must be square")
        self.n_total = A.shape[1]  # a rank of a sharded run may hold only its block of rows (rows x N)
        if len(A.indices) >= 2 ** 31:
            raise ValueError("graphs with >= 2^31 stored edges are not supported")
        data = A.data if A.data.dtype in (np.float32, np.float64) else A.data.astype(np.float64)
        if shard is not None and shard[0].world > 1:
            # Each rank ingests only its own block of the caller's rows (1 / world of the host -> device
            # traffic); the blocks are then all-gathered over NVLink so that every rank can compute the
            # same cell order and cut its shard of the reordered graph.
            indptr, indices, data = self._gather_blocks(A, data, shard[0])
            main = side = torch.cuda.current_stream() if indptr.is_cuda else None
            pending_data = None
        else:
            indptr = _to_dev(A.indptr, torch.int32)
            dev = indptr.device
            big = A.indices.dtype == np.int32 and A.indices.nbytes >= (8 << 20)
            # column indices: staged through the library's page-locked ring when the buffer is pageable
            # (several host threads, 4 MB chunks: PCIe speed instead of the driver's single-threaded staging)
            indices = _lib.HostUpload(A.indices, dev).wait() if big else _to_dev(A.indices, torch.int32)
            # the edge weights (2/3 of the bytes) are not needed by the ordering: they travel on a side
            # stream, staged by a library thread, while the breadth-first sweeps run
            main, side = torch.cuda.current_stream(), _upload_stream(dev)
            with torch.cuda.stream(side):
                pending_data = _lib.HostUpload(data, dev, background=True) if big else None
                data = pending_data.tensor if big else _to_dev(data)
            data.record_stream(main)
        self.order = self.inv = None
        mark("graph: indices queued for upload")
        if (want_reorder(self.n_total) if reorder is None else reorder) and self.n_total > 1:
            res = cuthill_mckee_order(indptr, indices, self.n_total, far_root=resident)
            mark("graph: cell order computed")
            if res is not None:
                self.order, self.inv = res
                deg = (indptr[1:] - indptr[:-1])[self.order]
                new_indptr = torch.zeros(self.n_total + 1, dtype=torch.int32, device=indptr.device)
                new_indptr[1:] = torch.cumsum(deg, 0)
                new_indices, new_data = torch.empty_like(indices), torch.empty_like(data)
                if pending_data is not None:
                    pending_data.wait()  # every chunk of the edge weights is queued on the side stream
                    pending_data = None
                if main is not side:
                    main.wait_stream(side)
                _lib.permute_csr(indptr, indices, data, self.order, self.inv, new_indptr, new_indices, new_data)
                indptr, indices, data = new_indptr, new_indices, new_data
        if pending_data is not None:
            pending_data.wait()
        if main is not side:
            main.wait_stream(side)
        self.halo_ids = None
        if shard is None:
            self.comm, self.row0, self.rows_per = None, 0, self.n_total
            self.indices_global = indices
        else:
            self.comm, self.row0, row1, self.rows_per = shard
            e0, e1 = int(indptr[self.row0].item()), int(indptr[row1].item())
            indptr = (indptr[self.row0:row1 + 1] - e0).contiguous()
            indices, data = indices[e0:e1].clone(), data[e0:e1].clone()
            self.indices_global = indices  # column sums / edge normalisation use global ids
            indices = self._plan_halo(indices, row1)
        self.n = indptr.numel() - 1  # rows held by this device
        self.nnz = int(indices.numel())
        self.indptr, self.indices, self.data = indptr, indices, data
        self._scaled = {}

    @staticmethod
    def _gather_blocks(A, data, comm):
        """Upload rows [o0, o1) of the host CSR (this rank's block of the caller's row order) and
        all-gather the blocks: returns the full (indptr, indices, data) on this rank's device."""
        n_total = A.shape[1]
        rows_per = (n_total + comm.world - 1) // comm.world
        o0 = min(comm.rank * rows_per, n_total)
        o1 = min(o0 + rows_per, n_total)
        if A.shape[0] != n_total:  # the host holds this rank's block only
            if A.shape[0] != o1 - o0:
                raise ValueError(f"rank {comm.rank} expects rows [{o0}, {o1}) of the graph, got {A.shape[0]} rows")
            o0, o1 = 0, A.shape[0]
        e0, e1 = int(A.indptr[o0]), int(A.indptr[o1])
        deg = np.zeros(rows_per, dtype=np.int32)
        deg[: o1 - o0] = np.diff(A.indptr[o0:o1 + 1])
        deg_all = comm.all_gather_rows(_to_dev(deg))[:n_total]
        indptr = torch.zeros(n_total + 1, dtype=torch.int32, device=deg_all.device)
        indptr[1:] = torch.cumsum(deg_all, 0)
        indices = torch.cat(comm.all_gather_padded(_to_dev(A.indices[e0:e1], torch.int32)))
        data = torch.cat(comm.all_gather_padded(_to_dev(data[e0:e1])))
        mark("graph: blocks gathered")
        return indptr, indices, data

    def _plan_halo(self, indices, row1):
        """kNN halo of this shard, gathered once: the sorted remote row ids its edges reference
        (``halo_ids``), the column ids renamed to positions in [own rows (rows_per slots) | halo rows],
        and who sends what (``send_idx`` / ``send_splits`` / ``recv_splits``)."""
        comm, r0, rows_per = self.comm, self.row0, self.rows_per
        cols = indices.long()
        local = (cols >= r0) & (cols < row1)
        halo = torch.unique(cols[~local])  # sorted => grouped by owner rank, ascending
        self.halo_ids = halo
        renamed = torch.where(local, cols - r0, rows_per + torch.searchsorted(halo, cols))
        owner = torch.div(halo, rows_per, rounding_mode="floor")
        self.recv_splits = torch.bincount(owner, minlength=comm.world).tolist()
        # every rank publishes its request list; a rank serves the ids that fall in its row block
        send_idx, self.send_splits = [], []
        for peer, wanted in enumerate(comm.all_gather_padded(halo)):
            mine = wanted[(wanted >= r0) & (wanted < row1)] - r0 if peer != comm.rank else wanted[:0]
            send_idx.append(mine)
            self.send_splits.append(int(mine.numel()))
        self.send_idx = torch.cat(send_idx)
        return renamed.to(torch.int32)

    def exchange_halo(self, ext):
        """Fill the halo rows ``ext[rows_per:]`` of a [rows_per + n_halo, ld] state with the current
        values of their owners' rows (``ext[:rows_per]`` on the owning ranks)."""
        comm, rows_per = self.comm, self.rows_per
        # a persistent send buffer per state width: a fresh 10-200 MB tensor per step would be recorded on
        # NCCL's stream and could not be recycled by the allocator until the collective has run
        buf = getattr(self, "_send_buf", None)
        if buf is None or buf.shape != (self.send_idx.numel(), ext.shape[1]) or buf.dtype != ext.dtype:
            buf = self._send_buf = torch.empty((self.send_idx.numel(), ext.shape[1]), dtype=ext.dtype,
                                               device=ext.device)
        send = torch.index_select(ext[:rows_per], 0, self.send_idx, out=buf)
        recv = ext[rows_per:]

        def via_all_gather():  # backends without all_to_all (gloo in the tests)
            recv.copy_(comm.all_gather_rows(ext[:rows_per].contiguous()).index_select(0, self.halo_ids))

        comm.exchange_rows(send, self.send_splits, recv, self.recv_splits, fallback=via_all_gather)

    def permute(self, t):
        """Rows of ``t`` (one per cell, caller's order) -> stored order."""
        return t if self.order is None else t.index_select(0, self.order)

    def unpermute(self, t):
        """Rows of ``t`` (one per cell, stored order, all cells) -> the caller's order."""
        return t if self.inv is None else t.index_select(0, self.inv.long())

    def scaled(self, self_weight=1, dtype=torch.float32):
        key = (float(self_weight), dtype)
        if key not in self._scaled:
            colsum = torch.zeros(self.n_total, dtype=torch.float64, device=self.indptr.device)
            _lib.graph_colsum(self.indptr, self.indices_global, self.data, colsum)
            if self.comm is not None:  # column sums need every shard's rows (_nam.py:28)
                self.comm.all_reduce(colsum)
            vals = torch.empty(self.nnz, dtype=dtype, device=colsum.device)
            diag = torch.empty(self.n, dtype=dtype, device=colsum.device)
            _lib.graph_scale(self.indptr, self.indices_global, self.data, colsum, self_weight, vals, diag,
                             row_offset=self.row0)
            self._scaled[key] = (vals, diag)
        return self._scaled[key]


class TilePlan:
    """Plan of the shared-memory-staged diffusion step (csrc/diffuse_tiled.cu) for one resident CSR:
    consecutive output rows are cut into tiles of at most ``tile_rows`` rows whose edges reference at
    most ``tile_sources`` distinct source rows (a tile that would exceed it is halved until it fits).

      tile_row [T + 1] int32   first output row of each tile
      tile_u   [T + 1] int32   offsets into ``usrc`` (multiples of 4)
      usrc     int32           sorted distinct source rows of every tile, padded to a multiple of 4
      epair    [nnz, 2] int32  per stored edge (CSR order): 128 * (position of its source in the tile's
                               list), bits of the fp32 weight

    Built with a handful of device sorts when the graph is made resident (not on the timed path)."""

    def __init__(self, indptr, indices, vals, n_src_rows):
        tile_rows, cap = _lib.diffuse_tile_limits()
        dev = indptr.device
        n = indptr.numel() - 1
        deg = (indptr[1:] - indptr[:-1]).long()
        row_of_edge = torch.repeat_interleave(torch.arange(n, device=dev), deg)
        cols = indices.long()
        starts = torch.arange(0, max(n, 1), tile_rows, device=dev)
        for _ in range(8):
            tile_of_row = torch.searchsorted(starts, torch.arange(n, device=dev), right=True) - 1
            keys = tile_of_row[row_of_edge] * n_src_rows + cols
            ukeys, inv = torch.unique(keys, sorted=True, return_inverse=True)
            utile = torch.div(ukeys, n_src_rows, rounding_mode="floor")
            ucount = torch.bincount(utile, minlength=starts.numel())
            bad = torch.nonzero(ucount > cap).reshape(-1)
            if bad.numel() == 0:
                break
            ends = torch.cat([starts[1:], torch.tensor([n], device=dev)])
            mid = (starts[bad] + ends[bad]) // 2
            if (mid == starts[bad]).any():
                raise _lib.CnaError(f"a single row references more than {cap} distinct rows")
            starts = torch.sort(torch.cat([starts, mid])).values
        else:
            raise _lib.CnaError("could not cut the graph into tiles")
        T = starts.numel()
        padded = (ucount + 3) // 4 * 4
        tile_u = torch.zeros(T + 1, dtype=torch.int64, device=dev)
        tile_u[1:] = torch.cumsum(padded, 0)
        first = torch.zeros(T + 1, dtype=torch.int64, device=dev)  # offsets into the unpadded unique list
        first[1:] = torch.cumsum(ucount, 0)
        usrc = torch.empty(int(tile_u[-1].item()), dtype=torch.int64, device=dev)
        # padding entries repeat the tile's first source row (any valid row will do: never referenced)
        usrc[:] = torch.repeat_interleave((ukeys % n_src_rows)[first[:-1].clamp(max=max(ukeys.numel() - 1, 0))],
                                          padded) if ukeys.numel() else 0
        pos_in_tile = torch.arange(ukeys.numel(), device=dev) - first[utile]
        usrc[tile_u[utile] + pos_in_tile] = ukeys % n_src_rows
        lpos = inv - first[tile_of_row[row_of_edge]]
        self.epair = torch.stack([(lpos * 128).to(torch.int32), vals.view(torch.int32)], dim=1).contiguous()
        self.tile_row = torch.cat([starts, torch.tensor([n], device=dev)]).to(torch.int32)
        self.tile_u = tile_u.to(torch.int32)
        self.usrc = usrc.to(torch.int32)
        self.n_tiles = T
        self.n_sources = int(ucount.sum().item())
        self.nnz = int(cols.numel())


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
    return DeviceGraph(get_connectivity(data), resident=False)  # one-shot: built for this call only


def _fingerprint(raw):
    """Cheap content fingerprint of a column buffer: a strided sample of 4096 values (microseconds; a full
    pass over a million ids would cost 0.5 ms of every resident call).  It catches a re-assigned column
    that landed on a reused address and any bulk in-place edit; a handful of ids edited in place between
    two calls on the same resident handle can escape it — call ``to_device`` again after such an edit."""
    step = max(1, raw.shape[0] // 4096)
    sample = raw[::step]
    if raw.dtype.kind in "iubf":
        return hash(sample.tobytes())
    return hash(tuple(sample.tolist()))


def sample_codes(data, sid_name):
    """Column order of ``pd.get_dummies(data.obs[sid_name])`` (``_nam.py:51``): the categories of a
    categorical column, otherwise the sorted unique values.  Returns (labels Index, int32 codes
    tensor on the device, cells-per-sample counts as float64 numpy)."""
    cache = getattr(data, "_codes", None)
    sid = data.obs[sid_name]
    categorical = isinstance(sid.dtype, pd.CategoricalDtype)
    raw = sid.cat.codes.to_numpy() if categorical else sid.to_numpy()
    # cache key for resident data: same column buffer and same content fingerprint => same codes (obs is
    # shared with the host object; an in-place edit keeps the buffer, a re-added column may reuse its address)
    key = (sid_name, raw.__array_interface__["data"][0], raw.shape[0], str(raw.dtype), _fingerprint(raw),
           hash(tuple(sid.cat.categories)) if categorical else None)
    if cache is not None and key in cache:
        return cache[key]
    if categorical:
        labels = pd.Index(sid.cat.categories)
        codes = raw
        if (codes < 0).any():
            raise ValueError(f"data.obs['{sid_name}'] contains missing values")
    elif raw.dtype.kind in "iu" and raw.dtype.itemsize <= 8 and len(raw) >= 100_000:
        # sorted unique values + inverse on the device (pd.factorize(sort=True) of 1M ids costs ~7 ms
        # of host time; integer ids have no NaN / object semantics to preserve)
        t = _to_dev(raw if raw.dtype != np.uint64 else raw.astype(np.int64))
        uniq, inverse, cnt = torch.unique(t, sorted=True, return_inverse=True, return_counts=True)
        labels = pd.Index(uniq.cpu().numpy().astype(raw.dtype, copy=False))
        out = (labels, inverse.to(torch.int32), cnt.cpu().numpy().astype(np.float64))
        if cache is not None:
            cache.clear()
            cache[key] = out
        return out
    else:
        codes, labels = pd.factorize(sid, sort=True)
        labels = pd.Index(labels)
    counts = np.bincount(codes, minlength=len(labels)).astype(np.float64)
    out = (labels, _to_dev(codes.astype(np.int32)), counts)
    if cache is not None:
        cache.clear()
        cache[key] = out
    return out
