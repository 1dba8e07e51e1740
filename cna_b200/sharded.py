"""Cell-axis sharding of the hot path across the GPUs of one box.

One process per GPU (``torchrun``); rank g owns a contiguous block of cells (of the stored,
Cuthill-McKee ordered graph): its rows of the kNN CSR, of the diffusion state and of the
residualised NAM.  Everything
sample-sized is replicated.  The data-path exchanges are

  * all-reduce of the graph's column sums (once per graph),
  * the kNN halo: the set of remote rows a shard's edges reference is computed once when the graph
    is made resident (``DeviceGraph``: column ids are renamed to [own rows | halo rows], the
    request lists are exchanged once); before every diffusion step after the first only those rows
    travel (one all_to_all_single).  In the stored Cuthill-McKee cell order a shard's halo is about as
    large as the shard itself instead of ~every other row,
  * all-gathers of one float64 per cell for the global medians (auto-stop, QC, ridge loop),
  * all-reduce of the n x n Gram, of max|ncorr| and of the T-bin null / observed histograms,
  * a broadcast of the permutation indices from rank 0 (the legacy RNG stream is drawn once),
  * an all-gather of the per-cell outputs, so every rank ends with the full ``data.obs`` columns.

``Comm`` is backend-agnostic (NCCL on GPUs; gloo for the CPU tests of the host logic).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_total, world, rank):
    """Contiguous row blocks of equal allocation ``rows_per = ceil(N / world)``; only the last
    non-empty block may be short.  Returns (row0, row1, rows_per)."""
    rows_per = (n_total + world - 1) // world
    r0 = min(rank * rows_per, n_total)
    r1 = min(r0 + rows_per, n_total)
    return r0, r1, rows_per


def slice_csr(A, r0, r1):
    """Rows [r0, r1) of a scipy CSR as (indptr rebased to 0, indices, data) without copying more
    than the slice."""
    A = A.tocsr()
    e0, e1 = int(A.indptr[r0]), int(A.indptr[r1])
    indptr = (A.indptr[r0:r1 + 1] - A.indptr[r0]).astype(np.int32)
    return indptr, A.indices[e0:e1], A.data[e0:e1]


class Comm:
    """Thin wrapper over a torch.distributed process group."""

    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.backend = dist.get_backend(group)

    def all_reduce(self, t, op="sum"):
        dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX, group=self.group)
        return t

    def all_gather_rows(self, local):
        """local: [rows_per, ...] on every rank -> [world * rows_per, ...] (rank-major)."""
        local = local.contiguous()
        out = torch.empty((self.world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                          device=local.device)
        if self.backend == "nccl":
            dist.all_gather_into_tensor(out, local, group=self.group)
        else:
            dist.all_gather(list(out.chunk(self.world, dim=0)), local, group=self.group)
        return out

    def all_gather_padded(self, t):
        """1-D tensors of different lengths -> list of per-rank tensors (one padded all-gather)."""
        n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
        sizes = self.all_gather_rows(n).tolist()
        pad = torch.zeros(max(max(sizes), 1), dtype=t.dtype, device=t.device)
        pad[: t.numel()] = t
        allp = self.all_gather_rows(pad).reshape(self.world, -1)
        return [allp[r, : sizes[r]] for r in range(self.world)]

    def exchange_rows(self, send, send_splits, recv, recv_splits, fallback=None):
        """Variable all-to-all of rows: ``send`` holds the rows for rank 0, 1, ... back to back
        (``send_splits`` rows each); ``recv`` receives ``recv_splits`` rows from each rank in rank
        order.  NCCL: one all_to_all_single over NVLink.  Other backends (the gloo tests) call
        ``fallback()``, which must fill ``recv`` by other means."""
        if self.backend == "nccl":
            dist.all_to_all_single(recv, send, recv_splits, send_splits, group=self.group)
        else:
            fallback()
        return recv

    def broadcast(self, t, src=0):
        dist.broadcast(t, src=src, group=self.group)
        return t

    def barrier(self):
        dist.barrier(group=self.group)


class ShardedData:
    """This rank's shard of an AnnData-like object, graph resident on its GPU.  Accepted by
    ``cna_b200.tl.association`` / ``nam`` in place of ``data``; results are written to the full
    ``data.obs`` on every rank."""

    def __init__(self, data, comm=None, resident=True):
        """``resident=False``: a shard built for one call (cheaper cell ordering: no pseudo-peripheral
        root sweep), like the graph ``association(data)`` builds for a host object."""
        from .tl._graph import DeviceGraph, get_connectivity
        self._host = data
        self.comm = comm or Comm()
        A = get_connectivity(data)
        n_total = A.shape[1]  # the host object may hold only this rank's block of rows (rows x N)
        r0, r1, rows_per = shard_bounds(n_total, self.comm.world, self.comm.rank)
        self.graph = DeviceGraph(A, shard=(self.comm, r0, r1, rows_per), resident=resident)
        self.obsp = getattr(data, "obsp", None)
        self.uns = getattr(data, "uns", None)
        self._codes = {}

    @property
    def obs(self):
        return self._host.obs


def shard_to_device(data, comm=None, resident=True):
    return data if isinstance(data, ShardedData) else ShardedData(data, comm, resident=resident)
