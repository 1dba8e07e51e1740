"""NAM construction, QC, residualisation and SVD — host side of kernels (i), (ii), (iii).

Mirrors ``src/cna/tools/_nam.py`` of the reference (same public names, arguments, defaults and
return types) with the arithmetic running on the GPU through the C-ABI in ``include/cna_b200.h``.
The device state is CELLS x SAMPLES (the transpose of the reference's DataFrames) in fp32.
"""
import os
from argparse import Namespace

import numpy as np
import pandas as pd
import torch

from .. import _lib
from ._graph import get_connectivity, graph_of, sample_codes, device, _to_dev  # noqa: F401
from ._out import select_output
from ._timing import mark

DEFAULT_RIDGES = [1e5, 1e4, 1e3, 1e2, 1e1, 1e0, 1e-1, 1e-2, 1e-3, 1e-4, 0]


def _round_up(x, m):
    return (x + m - 1) // m * m


_SKIP_BITS = 0x7FF4DEADBEEF0001  # CNA_MEDIAN_SKIP_BITS: entries cna_median_f64 ignores


def median_device(t, valid=None, comm=None, rows_per=None):
    """``np.median`` of a 1-D float64 device tensor, left on the device: returns a float64 tensor
    [2] = (median, number of entries considered).  The median is the mean of the two middle order
    statistics, NaN if any (valid) element is NaN or there is none.  ``valid`` (bool/uint8 mask)
    restricts it to a subset; with a ``comm`` it is over the concatenation of every rank's values (one
    all-gather of a padded vector in which masked and padding entries carry the skip pattern).
    Asynchronous: ``cna_median_f64`` (radix select), no sort, nothing copied back."""
    if t.dtype != torch.float64:
        t = t.double()
    t = t.contiguous()
    out = torch.empty(2, dtype=torch.float64, device=t.device)
    if comm is not None:
        bits = t.view(torch.int64)
        if valid is not None:
            bits = torch.where(valid.bool(), bits, torch.full_like(bits, _SKIP_BITS))
        pad = torch.full((rows_per,), _SKIP_BITS, dtype=torch.int64, device=t.device)
        pad[: t.numel()] = bits
        t, valid = comm.all_gather_rows(pad).view(torch.float64), None
    elif valid is not None:
        valid = valid.to(torch.uint8).contiguous()
    _lib.median(t, valid, out)
    return out


def device_median(t, valid=None, comm=None, rows_per=None):
    """``median_device`` copied back: a Python float (one synchronisation)."""
    return float(median_device(t, valid=valid, comm=comm, rows_per=rows_per)[0].item())


# ---------------------------------------------------------------------------------------------
# diffusion
# ---------------------------------------------------------------------------------------------
def diffuse_stepwise(data, s, maxnsteps=15, show_progress=False, self_weight=1):
    """``_nam.py:21-34``.  Generator over the state after each step.  ``s`` may be a numpy array or
    DataFrame (cells x k, float64 arithmetic on the GPU, numpy arrays are yielded) or a CUDA tensor
    (float32 or float64, CUDA tensors are yielded)."""
    out = select_output(show_progress)
    g = graph_of(data)
    if g.comm is not None:
        raise NotImplementedError("diffuse / diffuse_stepwise take the whole graph, not a cell-axis shard")
    on_device = torch.is_tensor(s)
    if on_device:
        cur = s if s.is_cuda else s.to(device())
        dtype = cur.dtype if cur.dtype in (torch.float32, torch.float64) else torch.float64
        cur = cur.to(dtype)
    else:
        frame = s if isinstance(s, pd.DataFrame) else None
        arr = np.asarray(s, dtype=np.float64)
        cur = _to_dev(arr)
        dtype = torch.float64
    if cur.dim() != 2 or cur.shape[0] != g.n:
        raise ValueError("s must be a 2-D array with one row per cell")  # colsums[:,None] needs 2-D
    cur = g.permute(cur.contiguous())
    vals, diag = g.scaled(self_weight, dtype)
    for i in range(maxnsteps):
        print("\ttaking step", i + 1, file=out)
        nxt = torch.empty_like(cur)
        _lib.diffuse_step(g.indptr, g.indices, vals, diag, cur, nxt, cur.shape[1])
        cur = nxt
        if on_device:
            yield g.unpermute(cur)
        else:
            res = g.unpermute(cur).cpu().numpy()
            yield pd.DataFrame(res, index=frame.index, columns=frame.columns) if frame is not None else res


def diffuse(data, s, nsteps, show_progress=False, self_weight=1):
    """``_nam.py:36-41``."""
    for s in diffuse_stepwise(data, s, maxnsteps=nsteps, show_progress=show_progress,
                              self_weight=self_weight):
        pass
    return s


class NamState:
    """Device-resident raw NAM: ``s`` [N x ld] fp32 (un-normalised diffusion state), the per-sample
    scaling 1/C (``_nam.py:73``), sample labels, the QC keep mask and step diagnostics."""

    def __init__(self, s, n_samples, labels, counts, cell_index):
        self.s = s
        self.S = n_samples
        self.labels = labels
        self.counts = counts
        with np.errstate(divide="ignore"):
            self.inv_count = _to_dev(1.0 / counts)
        self.cell_index = cell_index
        self._keep = None         # explicit uint8 keep mask (tests, callers that bring their own)
        self.qc_median = None     # device [2]: median of qc_kurt (cna_median_f64); None = no QC, keep all
        self.medkurt = []
        self.nsteps = 0
        self.comm = None  # set for cell-axis shards (cna_b200.sharded)
        self.row0 = 0
        self.rows_per = s.shape[0]
        self.graph = None  # DeviceGraph whose (possibly reordered) row order the state follows
        self.qc_kurt = None  # per-cell batch kurtosis when the last diffusion step produced it
        self.qc_batches = None

    @property
    def N(self):
        return self.s.shape[0]

    @property
    def qc_threshold(self):
        """``_nam.py:94``: max(6, 2 * median batch kurtosis) (Python's max: a NaN median gives 6); None
        without a QC.  Reads the median back from the device."""
        if self.qc_median is None:
            return None
        return max(6, 2 * float(self.qc_median[0].item()))

    @property
    def keep(self):
        """uint8 device mask of the cells that pass the QC (``_nam.py:96``), or None when every cell is
        kept.  The hot path never materialises it: the residualisation pass takes the decision per row
        from ``qc_kurt`` and the device-resident median."""
        if self._keep is not None or self.qc_median is None:
            return self._keep
        two_med = 2 * self.qc_median[0]
        thr = torch.where(two_med > 6, two_med, torch.full_like(two_med, 6.0))
        return (self.qc_kurt < thr).to(torch.uint8)  # NaN -> dropped

    @keep.setter
    def keep(self, mask):
        self._keep = mask


def _r2_p20(s, old, S):
    # print-only diagnostic of _nam.py:47-49,60,63 (runs only under show_progress)
    a, b = s[:, :S].double(), old[:, :S].double()
    r = ((a - a.mean(0)) * (b - b.mean(0))).mean(0) / a.std(0, unbiased=False) / b.std(0, unbiased=False)
    r2 = (r ** 2).cpu().numpy()
    return np.percentile(r2, 20)


_QC_PLANS = {}


def _series_key(ser):
    """Content key of a small pandas Series (values and index)."""
    v, ix = ser.to_numpy(), ser.index.to_numpy()
    return (v.tobytes() if v.dtype != object else tuple(v.tolist()), str(v.dtype),
            ix.tobytes() if ix.dtype != object else tuple(ix.tolist()), str(ix.dtype))


def _qc_plan(batches, labels, ld, dev):
    """Device tables for the QC statistic fused into the last diffusion step: batch of each sample
    column (int8, -1 for padding) and 1 / samples per batch; None when the fused kernel does not
    apply (one batch: no QC, _nam.py:89; more than 8 batches: separate kernel)."""
    if batches is None:
        return None
    # the plan depends on the batch Series and the sample labels only: kept per (content, labels, device)
    # — ~0.25 ms of pandas / numpy / two small uploads per call otherwise
    try:
        key = (_series_key(batches), id(labels), ld, dev.index)
    except Exception:  # noqa: BLE001 - unhashable content: no caching
        key = None
    if key is not None and key in _QC_PLANS:
        return _QC_PLANS[key][1]
    plan = None
    if len(np.unique(batches)) > 1:
        ub, order, off = _batch_segments(batches.reindex(labels).to_numpy())  # _nam.py:79
        if 2 <= len(ub) <= 8 and ld // 4 <= 128:  # the fused kernel's limits: <= 8 batches, <= 512 columns
            col_batch = np.full(ld, -1, dtype=np.int8)
            for b in range(len(ub)):
                col_batch[order[off[b]:off[b + 1]]] = b
            plan = _to_dev(col_batch), _to_dev(1.0 / np.diff(off).astype(np.float64))
    if key is not None:
        if len(_QC_PLANS) >= 8:
            _QC_PLANS.clear()
        _QC_PLANS[key] = (labels, plan)  # (the labels object is kept alive so that its id stays unique)
    return plan


def _nam_device(data, sid_name, nsteps=None, maxnsteps=15, self_weight=1, show_progress=False, codes=None,
                qc_batches=None):
    """``_nam.py:44-76`` on the device.  The first step never materialises the one-hot matrix.
    ``codes`` = an already computed ``sample_codes(data, sid_name)``; ``qc_batches`` = the per-sample
    batch Series of the QC that will follow (``_qc_device``): with a fixed ``nsteps`` its per-cell
    statistic is then produced by the last diffusion step itself."""
    out = select_output(show_progress)
    g = graph_of(data)
    labels, codes, counts = codes if codes is not None else sample_codes(data, sid_name)
    S = len(labels)
    ld = _round_up(S, 8)
    dev = codes.device
    comm = g.comm
    if codes.numel() != g.n_total:
        raise ValueError("data.obs and the connectivities graph disagree on the number of cells")
    vals, diag = g.scaled(self_weight, torch.float32)
    # a shard allocates rows_per slots for its own rows (equal on every rank) followed by its kNN
    # halo: the edges' column ids were renamed to positions in this buffer when the graph was made
    # resident, so a step is a purely local SpMM once the halo rows have been refreshed
    n_halo = 0 if comm is None else int(g.halo_ids.numel())
    # On one device every row of both buffers is written before it is read (the one-hot step writes all
    # ld columns, the later steps whole float4 groups computed from them; columns past the last group
    # are never touched by any consumer), so the 2 x 4*N*ld bytes are not zero-filled.  A shard also
    # holds padding rows up to rows_per that only the collectives see: those buffers start as zeros.
    alloc = torch.empty if comm is None else torch.zeros
    cur = alloc((g.rows_per + n_halo, ld), dtype=torch.float32, device=dev)
    nxt = alloc((g.rows_per + n_halo, ld), dtype=torch.float32, device=dev)
    codes = g.permute(codes)  # device rows follow the graph's stored cell order
    if comm is not None:
        own = torch.zeros(g.rows_per, dtype=codes.dtype, device=dev)
        own[: g.n] = codes[g.row0: g.row0 + g.n]
        codes = torch.cat([own, codes.index_select(0, g.halo_ids)])
    st = NamState(cur[: g.n], S, labels, counts, data.obs.index)
    st.comm, st.row0, st.rows_per, st.graph = comm, g.row0, g.rows_per, g
    need_stats = nsteps is None or show_progress
    kurt = torch.empty(g.n, dtype=torch.float64, device=dev) if need_stats else None
    qc = _qc_plan(qc_batches, labels, ld, dev) if (nsteps is not None and nsteps >= 2) else None
    if qc is not None:
        inv_ld = torch.zeros(ld, dtype=torch.float64, device=dev)
        inv_ld[:S] = st.inv_count
    old = None
    prevmedkurt = np.inf
    i = 0
    for i in range(maxnsteps):
        print("\ttaking step", i + 1, file=out)
        if i == 0:
            _lib.diffuse_onehot(g.indptr, g.indices, vals, diag, codes, S, cur, n_rows=g.n, row_offset=0)
        else:
            if show_progress:
                old = cur.clone()
            if comm is not None:
                g.exchange_halo(cur)  # only the rows other shards' edges reference travel
            if qc is not None and i + 1 == nsteps:  # last step: QC statistic straight from the accumulators
                st.qc_kurt = torch.empty(g.n, dtype=torch.float64, device=dev)
                st.qc_batches = qc_batches.reindex(labels)
                _lib.diffuse_step_qc(g.indptr, g.indices, vals, diag, cur, nxt, S, qc[0], inv_ld, qc[1],
                                     st.qc_kurt, n_rows=g.n, row_offset=0)
            else:
                _lib.diffuse_step(g.indptr, g.indices, vals, diag, cur, nxt, S, n_rows=g.n, row_offset=0)
            cur, nxt = nxt, cur
        if need_stats:
            _lib.row_kurtosis(cur[: g.n], S, st.inv_count, kurt)
            medkurt = device_median(kurt, comm=comm, rows_per=g.rows_per)  # _nam.py:59
            st.medkurt.append(medkurt + 3)
            print("\tmedian kurtosis:", medkurt + 3, file=out)
            if show_progress:
                r2 = _r2_p20(cur[: g.n], old[: g.n], S) if old is not None and comm is None else float("nan")
                print("\t20th percentile R2(t,t-1):", r2, file=out)
        if nsteps is None:
            if prevmedkurt - medkurt < 3 and i + 1 >= 3:  # _nam.py:65
                print("stopping after", i + 1, "steps", file=out)
                break
            prevmedkurt = medkurt
        elif i + 1 == nsteps:  # _nam.py:69
            break
    st.s = cur[: g.n]
    st.nsteps = i + 1
    return st


def _batch_segments(batch_values):
    """Group positions 0..len-1 by batch value (sorted unique order, like ``np.unique`` at
    ``_nam.py:81``).  Returns (unique values, order int32, offsets int32)."""
    b = np.asarray(batch_values)
    ub, inv = np.unique(b, return_inverse=True)
    order = np.argsort(inv, kind="stable").astype(np.int32)
    off = np.zeros(len(ub) + 1, dtype=np.int32)
    np.cumsum(np.bincount(inv, minlength=len(ub)), out=off[1:])
    return ub, order, off


def _qc_device(st, batches, show_progress=False):
    """``_nam.py:85-99``.  Leaves the per-cell batch kurtosis (``st.qc_kurt``) and its median
    (``st.qc_median``) on the device; the keep decision ``kurt < max(6, 2 median)`` is taken by whoever
    consumes the rows (``cna_resid_pass``, ``st.keep``).  Nothing is copied back unless progress output
    is wanted."""
    out = select_output(show_progress)
    st._keep, st.qc_median = None, None
    if len(np.unique(batches)) == 1:  # _nam.py:89
        return
    b = batches.reindex(st.labels)  # _nam.py:79
    dev = st.s.device
    if not (st.qc_kurt is not None and st.qc_batches is not None and b.equals(st.qc_batches)):
        # not already produced by the last diffusion step
        ub, order, off = _batch_segments(b.to_numpy())
        st.qc_kurt = torch.empty(st.N, dtype=torch.float64, device=dev)
        _lib.batch_kurtosis(st.s, st.inv_count, _to_dev(order), _to_dev(off), st.qc_kurt)
    st.qc_median = median_device(st.qc_kurt, comm=st.comm, rows_per=st.rows_per)  # _nam.py:94
    if show_progress:
        print("throwing out neighborhoods with batch kurtosis >=", st.qc_threshold, file=out)
        nkeep = st.keep.sum().double().reshape(1)
        print("keeping", int((st.comm.all_reduce(nkeep) if st.comm is not None else nkeep).item()),
              "neighborhoods", file=out)


def nam(data, sid_name, batches=None, nsteps=None, self_weight=1, max_frac_pcs=0.15, suffix="",
        ks=None, show_progress=False, **kwargs):
    """``_nam.py:179-193``.  Returns (DataFrame samples x kept cells float64, keep mask bool[N])."""
    out = select_output(show_progress)
    if batches is None:  # _nam.py:185-186
        u = data.obs[sid_name].unique()
        batches = pd.Series(np.ones(len(u)), index=u)
    print("computing NAM", file=out)
    st = _nam_device(data, sid_name, nsteps=nsteps, self_weight=self_weight, show_progress=show_progress,
                     qc_batches=batches)
    _qc_device(st, batches, show_progress=show_progress)
    return nam_frame(st, sid_name), keep_mask(st)


def to_caller_order(st, t):
    """Per-cell device tensor in the state's row order -> the caller's cell order (all cells)."""
    if st.comm is not None:
        raise NotImplementedError("per-cell matrices of a cell-axis shard are not gathered")
    return t if st.graph is None else st.graph.unpermute(t)


def keep_mask(st):
    keep = st.keep
    if keep is None:
        return np.repeat(True, st.N)
    return to_caller_order(st, keep).bool().cpu().numpy()


def nam_frame(st, sid_name, rows=None, cols=None):
    """Materialise (a part of) the QC'd NAM as the reference's samples x cells DataFrame."""
    x = to_caller_order(st, st.s)[:, :st.S].double() * st.inv_count  # _nam.py:73
    keep = keep_mask(st)
    if not keep.all():
        x = x[torch.as_tensor(keep, device=x.device)]
    arr = x.t().contiguous().cpu().numpy()
    df = pd.DataFrame(arr, index=st.labels, columns=st.cell_index[keep], dtype=float)
    df.index.name = sid_name  # _nam.py:74
    return df


# ---------------------------------------------------------------------------------------------
# SVD of a NAM given on the host (public helper)
# ---------------------------------------------------------------------------------------------
def svd_nam(NAM):
    """``_nam.py:102-115``.  ``NAM`` is the reference's samples x cells DataFrame; the centring /
    standardisation and the Gram contraction run on the GPU, the n x n SVD on the host."""
    idx, cols = NAM.index, NAM.columns
    x64 = _to_dev(NAM.to_numpy(dtype=np.float64)).t().contiguous()  # cells x samples
    n = x64.shape[1]
    x64 = x64 - x64.mean(dim=1, keepdim=True)  # :103
    x64 = x64 / x64.std(dim=1, unbiased=True, keepdim=True)  # :104 pandas ddof=1
    ld = _round_up(n, 8)
    x = torch.zeros((x64.shape[0], ld), dtype=torch.float32, device=x64.device)
    x[:, :n] = x64
    del x64
    xp = planes_of(x, n)
    U, svs, _ = gram_svd(x, n, planes=xp)
    pcs = ["PC" + str(i) for i in range(1, n + 1)]
    V = nbhd_loadings(x, n, U, svs, planes=xp)
    return (pd.DataFrame(U, index=idx, columns=pcs), pd.Series(svs, index=pcs),
            pd.DataFrame(V, index=cols, columns=pcs))


TC_GRAM_MAX_N = 512  # four 128-row output tiles in groups of <= 512 TMEM columns


def planes_of(x, n):
    """fp16 hi/lo operand planes of the fp32 matrix x[:, :n] (tensor-core operand format)."""
    return _lib.split_f16(x, n)


# cna_sym_eig_top is used while the packed Gram fits in one SM's shared memory (n <= 208: 1.2 ms at n = 200
# against 2.1 ms + two PCIe hops for LAPACK); above that its matrix lives in L2 and LAPACK dsyevr with a few
# BLAS threads wins (n = 500: 22 ms on the device, ~13 ms on the host; measured on config E's per-rank
# share, 49.9 against 41.4 ms per call).  CNA_B200_HOST_EIG=1 forces the LAPACK route (cross-check),
# CNA_B200_DEVICE_EIG_MAX_N overrides the limit (the kernel itself takes n <= 512).
DEVICE_EIG_MAX_N = int(os.environ.get("CNA_B200_DEVICE_EIG_MAX_N", "208"))
if os.environ.get("CNA_B200_HOST_EIG"):
    DEVICE_EIG_MAX_N = 0


def gram_device(x, n, comm=None, planes=None):
    """NAM.NAM^T of ``_nam.py:105`` as an [n x n] float64 device tensor (tcgen05 kernel on the fp16
    hi/lo planes when n <= 256, CUDA cores otherwise; summed over shards when the cell axis is
    sharded).  Asynchronous: nothing is copied to the host."""
    dev = x.device if x is not None else planes.t.device
    G = torch.zeros((n, n), dtype=torch.float64, device=dev)
    if n <= TC_GRAM_MAX_N:
        _lib.gram_tc(planes if planes is not None else planes_of(x, n), n, G)
    else:
        _lib.gram(x, n, G)
    if comm is not None:
        comm.all_reduce(G)
    return G


_BLAS = None


def _single_threaded_blas():
    """Context manager capping the BLAS/LAPACK thread pool for the n x n decomposition on the host (the
    full-result surface and n > DEVICE_EIG_MAX_N): a few ms of work; with OpenBLAS's default (one spinning
    thread per core) it collapses to 10x that when something else is using the same cores.  One thread
    while the host permutation engine may be running, ``CNA_B200_BLAS_THREADS`` (default 4) otherwise."""
    global _BLAS
    try:
        if _BLAS is None:
            from threadpoolctl import ThreadpoolController
            _BLAS = ThreadpoolController()
        from . import _stats
        limit = int(os.environ.get("CNA_B200_BLAS_THREADS", "4")) if _stats._device_draw_enabled() else 1
        return _BLAS.limit(limits=max(1, limit), user_api="blas")
    except Exception:  # threadpoolctl missing: run with whatever the BLAS does by default
        import contextlib
        return contextlib.nullcontext()


def svd_of_gram(Gh, top=None):
    """``_nam.py:105``: U, svs, _ = np.linalg.svd(Gram) on the host (n x n).

    ``top`` = number of leading components the caller will actually read.  The association test
    uses U[:, :max(ks)] only (as the projector U_k U_k^T, so the sign of a column is irrelevant) and
    never the trailing ~85 % of the decomposition, so unless the full result surface is requested
    the symmetric PSD Gram is eigen-decomposed (numpy's LAPACK dsyevd, which — unlike scipy's f2py
    wrappers — releases the GIL, so it can run beside the thread that launches the null kernels)
    instead of a full dgesdd: same subspaces and singular values to ~1e-15 in less than half the
    time.  The returned arrays keep the full shapes (trailing columns / values are zero and must not
    be read)."""
    Gh = (Gh + Gh.T) / 2  # both triangles hold the same products up to summation order; keep it exact
    n = Gh.shape[0]
    if top is None or top >= n:
        with _single_threaded_blas():
            U, svs, _ = np.linalg.svd(Gh)
    else:
        # LAPACK dsyevr restricted to the leading `top` eigenpairs (relatively robust representations):
        # half the time of the full divide-and-conquer decomposition at n = 200
        import scipy.linalg as sl
        with _single_threaded_blas():  # spinning BLAS workers would starve the permutation draw's threads
            w, v = sl.eigh(Gh, subset_by_index=[n - top, n - 1], driver="evr")  # ascending
        U = np.zeros((n, n))
        svs = np.zeros(n)
        U[:, :top] = v[:, ::-1]
        svs[:top] = np.maximum(w[::-1], 0.0)
    mark("svd done")
    return U, svs, Gh


def gram_svd(x, n, comm=None, planes=None):
    G = gram_device(x, n, comm=comm, planes=planes)
    Gh = G.cpu().numpy()
    mark("gram on host")
    return svd_of_gram(Gh)


def nbhd_loadings(x, n, U, svs, rows=None, planes=None):
    """``_nam.py:106``: V = NAM^T U / sqrt(svs) (cells x n).  The trailing ~null-space columns divide
    by ~0 exactly like the reference.  The product X.U runs on the tensor cores (columns of U are
    unit vectors, inside fp16 range); the per-column 1/sqrt(svs) scaling is applied afterwards in
    float64 so that the near-null-space blow-up never passes through fp16."""
    xp = planes if planes is not None else planes_of(x, n)
    ldb = _round_up(n, 4)
    ut = torch.zeros((n, _round_up(n, 4)), dtype=torch.float32, device=xp.t.device)
    ut[:, :n] = _to_dev(np.ascontiguousarray(U.T, dtype=np.float32))
    utp = _lib.split_f16(ut, n)  # planes of U^T: row j = column j of U
    out = torch.empty((xp.rows, ldb), dtype=torch.float32, device=xp.t.device)
    _lib.right_multiply_tc(xp, n, utp, n, out)
    with np.errstate(divide="ignore", invalid="ignore"):
        scale = _to_dev(1.0 / np.sqrt(svs))
    V = out[:, :n]
    if rows is not None:  # a row selector: boolean mask / index tensor, or a callable
        V = rows(V) if callable(rows) else V[rows]
    return (V.double() * scale).cpu().numpy()


# ---------------------------------------------------------------------------------------------
# residualisation (host part: the small design-matrix algebra)
# ---------------------------------------------------------------------------------------------
def design_matrix(covs, batches, n):
    """``_nam.py:123-139``.  Returns (C [n x r] float64, number of batch columns).  pandas ``.std``
    is ddof=1."""
    if covs is None:
        cov = np.ones((n, 0))
    else:
        cov = np.asarray(covs, dtype=np.float64).reshape(n, -1)
        cov = (cov - cov.mean(axis=0)) / cov.std(axis=0, ddof=1)  # :126
    if batches is None or len(np.unique(batches)) == 1:  # :128
        return cov, 0
    ub = np.unique(batches)  # get_dummies column order, :137
    B = (np.asarray(batches)[:, None] == ub[None, :]).astype(np.float64)
    B = (B - B.mean(axis=0)) / B.std(axis=0, ddof=1)  # :138
    return np.concatenate([B, cov], axis=1), B.shape[1]


def projector(C, nb, ridge):
    """``_nam.py:145-146`` (or ``:133`` when there are no batch columns): W [r x n], M = I - C.W."""
    n, r = C.shape
    CtC = C.T.dot(C)
    if nb > 0:
        L = np.diag([1.0] * nb + [0.0] * (r - nb))
        CtC = CtC + ridge * n * L
    return np.linalg.solve(CtC, C.T)


def resid_nam_device(st, colmap, covs, batches, y_std, ridges=None, show_progress=False, want_x=True,
                     speculate=False, design=None, qc_in_pass=False):
    """``_nam.py:118-177`` + ``_association.py:178-185`` + ``:77`` on the device.

    colmap : int array, state column of each of the n selected samples (phenotype order)
    covs   : [n x c] array or None;  batches : length-n array;  y_std : standardised phenotype.
    Returns a Namespace with the device tensors ``x`` [N x ld] (rows of dropped cells are zero),
    ``ncorr`` [N], ``valid`` [N] and the host matrices M, C, W_last, r, ridge log.

    The ridge walk (:141-155) stops at the first ridge whose median batch kurtosis is <= 6.  With
    ``speculate`` only the first ridge is run and its median stays on the device (``res.ridge_median``):
    the caller queues the downstream kernels behind it, reads the median back together with their
    results and calls ``res.settle(median)``, which either confirms the guess or walks the remaining
    ridges (returning False: the downstream work must be repeated).  Without it the walk synchronises
    after every ridge, like the reference."""
    out = select_output(show_progress)
    n = len(colmap)
    dev = st.s.device
    C, nb = design if design is not None else design_matrix(covs, batches, n)
    r = C.shape[1]
    ld = _round_up(n, 8)
    # the fp32 matrix is only needed for the full result surface and for the CUDA-core Gram that
    # takes over beyond the tensor-core kernel's 512-sample limit
    want_x = want_x or n > TC_GRAM_MAX_N
    x = torch.empty((st.N, ld), dtype=torch.float32, device=dev) if want_x else None
    planes = _lib.Planes(st.N, n, dev)
    ncorr = torch.empty(st.N, dtype=torch.float64, device=dev)
    valid = torch.empty(st.N, dtype=torch.uint8, device=dev)
    colmap_d = _to_dev(np.asarray(colmap, dtype=np.int32))
    y_d = _to_dev(np.asarray(y_std, dtype=np.float64))
    res = Namespace(x=x, planes=planes, ncorr=ncorr, valid=valid, n=n, r=r, C=C, ridge_log=[], comm=st.comm,
                    ridge_median=None, settle=lambda med=None: True)
    C_d = _to_dev(C) if r else None
    row_keep = st._keep  # an explicit mask wins; otherwise the QC decision is taken inside the pass
    qc = Namespace(kurt=st.qc_kurt if (row_keep is None and st.qc_median is not None) else None,
                   pending=bool(qc_in_pass) and nb > 1 and row_keep is None and st.qc_median is None)

    def run(Wcum, seg=None, kurt=None):
        W_d = _to_dev(np.ascontiguousarray(Wcum)) if r else None
        qc_out = torch.empty(st.N, dtype=torch.float64, device=dev) if qc.pending else None
        _lib.resid_pass(st.s, st.inv_count, colmap_d, row_keep, C_d, W_d,
                        seg[0] if seg else None, seg[1] if seg else None, y_d, x, kurt, ncorr, valid,
                        planes=planes, qc_kurt=qc.kurt, qc_median=st.qc_median if qc.kurt is not None else None,
                        qc_out=qc_out)
        if qc_out is not None:
            # the QC statistic (_nam.py:78-82) came out of this pass: median, threshold and the late decision
            # (_nam.py:94-99) follow on the device; later ridge stages reuse it as an input
            qc.pending = False
            st.qc_kurt = qc.kurt = qc_out
            st.qc_median = median_device(qc_out, comm=st.comm, rows_per=st.rows_per)
            _lib.qc_fixup(qc_out, st.qc_median, x, planes, kurt, ncorr, valid)

    if nb == 0:
        if r > 0:  # _nam.py:133
            W = projector(C, 0, 0.0)
            res.M = np.eye(n) - C.dot(W)
        else:  # :131
            W = np.zeros((0, n))
            res.M = np.eye(n)
        res.W_last = W
        run(W)
        return res

    _, order, off = _batch_segments(batches)
    seg = (_to_dev(order), _to_dev(off))
    kurt = torch.empty(st.N, dtype=torch.float64, device=dev)
    ridge_list = list(DEFAULT_RIDGES if ridges is None else ridges)
    walk = Namespace(Wcum=np.zeros((r, n)), next=0)

    def step():
        """One ridge of :141-148; returns the device median of the batch kurtosis (:150-153)."""
        ridge = ridge_list[walk.next]
        walk.next += 1
        W = projector(C, nb, ridge)
        # the reference applies M = I - C.W cumulatively (:148); the product of such projectors is
        # again I - C.W' with W' = W_prev + W - (W C) W_prev
        walk.Wcum = walk.Wcum + W - W.dot(C).dot(walk.Wcum)
        run(walk.Wcum, seg, kurt)
        res.W_last = W
        res.M = np.eye(n) - C.dot(W)  # only the last M survives (:169)
        return ridge, median_device(kurt, valid=valid, comm=st.comm, rows_per=st.rows_per)

    def log(ridge, med):
        res.ridge_log.append((ridge, med))
        print("\twith ridge", ridge, "median batch kurtosis = ", med, file=out)
        return med <= 6 or walk.next >= len(ridge_list)  # :154-155

    def settle(med=None):
        """Confirm the speculative first ridge with its median (read back by the caller), or walk on."""
        res.settle = lambda med=None: True
        if med is None:
            med = float(res.ridge_median[0].item())
        if log(ridge_list[0], med):
            return True
        while True:
            ridge, med_d = step()
            if log(ridge, float(med_d[0].item())):
                return False

    if not ridge_list:
        raise UnboundLocalError("cannot access local variable 'M' where it is not associated with a value")
    ridge, res.ridge_median = step()
    if speculate and not show_progress:
        res.settle = settle
    else:
        settle()
    return res
