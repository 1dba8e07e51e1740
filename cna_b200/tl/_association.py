"""Association driver: global permutation test + neighbourhood-level FDRs.

Mirrors ``src/cna/tools/_association.py`` of the reference: same signature, side effects on
``data.obs``, exceptions, warnings and result Namespace.  The NAM never leaves the GPU unless
``return_full=True`` asks for the big matrices.
"""
import os
import warnings
from argparse import Namespace

import numpy as np
import pandas as pd
import scipy.stats as st
import torch

from .. import _lib
from . import _graph, _nam, _stats
from ._graph import _to_dev
from ._out import select_output
from ._timing import mark, report


def check_inputs(data, y, sid_name, batches, covs, donorids, allow_low_sample_size, present=None):
    """``_association.py:131-173`` — same checks, exception types and messages.  ``present`` may
    carry the already-known unique sample ids of ``data.obs[sid_name]``."""
    if not isinstance(y, pd.Series):
        raise TypeError(f"'y' must be a pandas Series, but got {type(y)}")
    if batches is not None and not isinstance(batches, pd.Series):
        raise TypeError(f"'batches' must be a pandas Series, but got {type(batches)}")
    if covs is not None and not isinstance(covs, pd.DataFrame):
        raise TypeError(f"'covs' must be a pandas DataFrame, but got {type(covs)}")
    if donorids is not None and not isinstance(donorids, pd.Series):
        raise TypeError(f"'donorids' must be a pandas Series, but got {type(donorids)}")
    if present is None:
        present = data.obs[sid_name].unique()
    if not y.index.isin(present).all():
        print("WARNING: index of 'y' contains values not present in 'data[sid_name]'. "
              "These samples will be ignored.")
    if not pd.Index(present).isin(y.index).all():
        raise ValueError("'data[sid_name]' contains values not present in the index of 'y'.")
    if batches is not None and donorids is not None:
        raise ValueError("We do not currently support conditioning on batch "
                         "while also accounting for multiple samples per donor")
    if batches is None:
        batches = pd.Series(np.ones(len(y)), index=y.index)
    if covs is not None:
        if covs.index.equals(y.index):  # same rows in the same order: the row-wise any() in numpy
            missing = y.isna().to_numpy() | covs.isna().to_numpy().any(axis=1)
            filter_samples = pd.Series(~missing & y.index.isin(present), index=y.index)
        else:  # pandas aligns the two indexes
            filter_samples = ~(y.isna() | covs.isna().any(axis=1)) & y.index.isin(present)
        if donorids is not None:
            print("WARNING: We currently do not account for multiple samples per donor "
                  "when conditioning on covariates. This conditioning may therefore account "
                  "only incompletely for the covariates of interest. We expect this to make "
                  "only minor differences in most cases, but we have not investigated it formally")
    else:
        filter_samples = ~np.isnan(y) & y.index.isin(present)
    if filter_samples.sum() < 10 and not allow_low_sample_size:
        raise ValueError(
            "You are supplying phenotype information on fewer than 10 samples. This may lead to "
            "poor power at low sample sizes because our null distribution is one in which each "
            "sample's single-cell profile is unchanged but the sample labels are randomly "
            "assigned. If you want to run an analysis at this sample size despite the possibility of low "
            "power, you can do so by invoking the association(...) function with the argument "
            "allow_low_sample_size=True.")
    return batches, filter_samples


_PINNED = {}


def _pinned(slot, shape, dtype):
    """A cached page-locked host tensor per ``slot`` (re-allocated when the shape changes).  The
    caller consumes it before the next call of the same slot."""
    key = (dtype, tuple(shape))
    held = _PINNED.get(slot)
    if held is None or held[0] != key:
        held = _PINNED[slot] = (key, torch.empty(shape, dtype=dtype, pin_memory=True))
    return held[1]


class _Readback:
    """Several small device tensors -> one page-locked staging buffer with one event: ``get()`` waits
    for the event only, not for whatever has been queued on the stream since."""

    def __init__(self, slot, tensors):
        self.views = []
        sizes = [(t.numel() * t.element_size() + 15) // 16 * 16 for t in tensors]  # every view 16-byte aligned
        buf = _pinned(slot, (max(sum(sizes), 16),), torch.uint8)
        off = 0
        for t, size in zip(tensors, sizes):
            nbytes = t.numel() * t.element_size()
            host = buf[off:off + nbytes].view(t.dtype).reshape(t.shape)
            host.copy_(t, non_blocking=True)
            self.views.append(host)
            off += size
        self.event = torch.cuda.Event()
        self.event.record(_lib.current_stream_object())

    def get(self):
        self.event.synchronize()
        return [v.numpy() for v in self.views]


_SIDE = {}


def _side_stream(dev, tag="copy"):
    """Side streams per device: "copy" for the copies of the per-cell result columns (they overlap the
    kernels queued after the pass that produced them), "eig" for the single-CTA eigensolver."""
    key = (dev.type, dev.index, tag)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev, priority=-1 if tag == "eig" else 0)
    return _SIDE[key]


_SMS = {}


def _sm_count(dev):
    key = (dev.type, dev.index)
    if key not in _SMS:
        _SMS[key] = torch.cuda.get_device_properties(dev).multi_processor_count
    return _SMS[key]


def _to_host_owned(t, stream=None):
    """Device tensor -> page-locked host tensor of its own, copied on ``stream`` (after everything
    queued so far on the current stream).  Returns (host tensor, event)."""
    buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    cur = _lib.current_stream_object()
    stream = stream or cur
    if stream is not cur:
        stream.wait_stream(cur)
    with torch.cuda.stream(stream):
        buf.copy_(t, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(stream)
    t.record_stream(stream)
    return buf, ev


def _adopt_column(obs, key, values):
    """``obs[key] = values`` without copying the N values a second time: the column adopts
    ``values`` (an array nothing else writes to; the ndarray keeps its buffer alive), where assigning
    the bare ndarray makes pandas copy it (~1 ms per million cells, on the critical path of a call)."""
    try:
        obs[key] = pd.Series(values, index=obs.index, copy=False)
    except Exception:  # noqa: BLE001 - an obs container that cannot adopt a Series gets the plain copy
        obs[key] = values


def _resid_tables_fit(n, r, nbk):
    """Mirror of the shared-memory budget of the linear-functional residualisation kernel
    (csrc/nam_pass.cu:cna_resid_pass, the only kernel that emits ``qc_out``), for n <= 256 samples."""
    nq = (n + 31) // 32
    if nq > 8:
        return False
    rows, warps, ldn = 4, 8, 32 * nq
    m1p = -(-(2 + r + nbk) // (16 // rows)) * (16 // rows)
    doubles = (m1p + r + 1) * ldn + nbk * r + r + 1 + warps * rows * (m1p + 3)
    return 8 * doubles + 4 * ldn <= 100 * 1024


def _all_samples_selected(labels, y, batches, covs):
    """True when the call selects every sample of the data, in label order, with complete phenotype / batch /
    covariate values, 2..16 batches, and a design small enough for the kernel that emits the statistic (cheap
    numpy checks: this runs before the first kernel is queued)."""
    try:
        if batches is None or len(y) != len(labels) or len(labels) > 256:
            return False
        if not (y.index.equals(labels) and (batches.index is y.index or batches.index.equals(y.index))):
            return False
        yv, bv = y.to_numpy(), batches.to_numpy()
        if yv.dtype.kind not in "fiub" or bv.dtype.kind not in "fiub":
            return False
        if (yv.dtype.kind == "f" and np.isnan(yv).any()) or (bv.dtype.kind == "f" and np.isnan(bv).any()):
            return False
        if covs is not None:
            cv = covs.to_numpy()
            if not (covs.index is y.index or covs.index.equals(y.index)) or cv.dtype.kind not in "fiub":
                return False
            if (cv.dtype.kind == "f" and np.isnan(cv).any()) or cv.ndim != 2:
                return False
        nb = len(np.unique(bv))
        ncov = 0 if covs is None else cv.shape[1]
        return 2 <= nb <= 16 and _resid_tables_fit(len(labels), nb + ncov, nb)
    except Exception:  # noqa: BLE001 - anything unusual takes the general route
        return False


def default_ks(n):
    """``_association.py:25-28``."""
    incr = max(int(0.02 * n), 1)
    maxnpcs = max(min(4 * incr, int(n / 5)), 1)
    return np.arange(incr, maxnpcs + 1, incr)


_POOL = None


def _f_sf(f, dfn, dfd):
    """``st.f.sf`` (:46).  scipy's incomplete-beta ufunc releases the GIL, so large inputs are
    evaluated in row chunks on a small thread pool (same function, same bits)."""
    global _POOL
    if f.shape[0] < 2048:
        # st.f.sf(f, dfn, dfd) without scipy.stats' argument machinery: same special function, same bits
        from scipy.special import fdtrc
        dfn, dfd = np.broadcast_arrays(np.asarray(dfn, dtype=np.float64), np.asarray(dfd, dtype=np.float64))
        with np.errstate(all="ignore"):
            p = fdtrc(dfn, dfd, f)
            p = np.where(f <= 0, np.where(np.isnan(f), np.nan, 1.0), p)
            return np.where((dfd > 0) & (dfn > 0), p, np.nan)
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        import os
        _POOL = ThreadPoolExecutor(max_workers=max(1, min(8, (os.cpu_count() or 2) // 2)))
    chunks = np.array_split(np.arange(f.shape[0]), _POOL._max_workers)
    parts = list(_POOL.map(lambda ix: st.f.sf(f[ix[0]:ix[-1] + 1], dfn, dfd), [c for c in chunks if len(c)]))
    return np.concatenate(parts, axis=0)


def _f_pvalues(ssered, ssefull, ks, n, r):
    """``_association.py:41-48`` vectorised over permutations: returns (p, r2), each [K x len(ks)]."""
    ks = np.asarray(ks, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        f = ((ssered[:, None] - ssefull) / ks) / (ssefull / n)  # :45 (divides by n, not dof)
        p = _f_sf(f, ks, n - (1 + r + ks))  # :46
        r2 = 1 - ssefull / ssered[:, None]  # :47
    return p, r2


def _pick(p, r2, ks):
    """``_association.py:60``: k_ = nanargmin(ps) per permutation."""
    with np.errstate(invalid="ignore"):
        pick = np.nanargmin(p, axis=1)
    rows = np.arange(p.shape[0])
    return np.asarray(ks)[pick], p[rows, pick], r2[rows, pick]


THRESHOLD_CAP = 512  # capacity of the device threshold tables (np.arange(m/4, m, m/400) has 300 or 301)

# diagnostics of the most recent association() call on this process (bench.py's `check` block):
# leading singular values of the residualised NAM, FDR thresholds, the k chosen, timings are not kept
LAST = Namespace()


class _Columns:
    """The two per-cell output columns (``_association.py:228-237``).  ``data.obs[key_added]`` is known
    as soon as the NAM is residualised and the per-cell FDR as soon as the null histograms are: both
    are copied back on a side stream, behind the kernels that produce them and beside everything queued
    afterwards, each into a page-locked buffer of its own that the obs column then adopts."""

    def __init__(self, data, key_added, stn, res, gather):
        self.data, self.key, self.stn, self.res, self.gather = data, key_added, stn, res, gather
        self.coef = self.fdr = None
        self.side = _side_stream(res.ncorr.device)

    def _full(self, t):
        """This rank's rows (stored order) -> all cells in the caller's order."""
        stn = self.stn
        if stn.comm is not None:
            pad = torch.zeros((stn.rows_per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            pad[: stn.N] = t
            t = stn.comm.all_gather_rows(pad)[: len(self.data.obs)]
        return t if stn.graph is None else stn.graph.unpermute(t)

    def start_coef(self):
        res = self.res
        coef = torch.where(res.valid.bool(), res.ncorr, torch.full_like(res.ncorr, float("nan")))
        if self.gather:
            self.coef = _to_host_owned(self._full(coef), self.side)

    def start_fdr(self, thr_d, pmin_d, count_d):
        """Per-cell FDR (:234-237) from the device-resident threshold table and running minimum of the FDR
        (``cna_fdr_table``): queued right behind the null GEMM, no host round trip.  (With
        local_test=False the reference writes the coefficients and then crashes looking up FDRs at :235;
        here the FDR column is simply not written.)"""
        res = self.res
        coef_d = torch.empty_like(res.ncorr)
        fdr_d = torch.empty_like(res.ncorr)
        _lib.cell_fdr_dev(res.ncorr, res.valid, thr_d, pmin_d, count_d, coef_d, fdr_d)
        if self.gather:
            self.fdr = _to_host_owned(self._full(fdr_d), self.side)

    def write(self):
        obs = self.data.obs
        if self.coef is not None:
            if self.key in obs:
                warnings.warn(f"Key '{self.key}' already exists in data.obs. Overwriting.")
            self.coef[1].synchronize()
            _adopt_column(obs, self.key, self.coef[0].numpy())
            mark("coef column written")
        if self.fdr is not None:
            self.fdr[1].synchronize()
            _adopt_column(obs, f"{self.key}_fdr", self.fdr[0].numpy())
        mark("obs written")


def _association(res, perms, Nnull=1000, local_test=True, show_progress=False, columns=None, host_draw=True):
    """``_nam.py:163`` (Gram + SVD of the residualised NAM) and ``_association.py:10-129`` after
    seeding / permutation drawing (done by the caller so that they overlap with the NAM kernels).
    ``res`` carries the device-resident residualised NAM (``res.planes`` / ``res.x``), M, r, the
    standardised phenotype and ks; U, svs and the Gram are added to it.

    The device work is queued in two phases, neither of which waits for the host:
      A  Gram, max |ncorr|, the FDR thresholds derived from it (on the device) and one packed read-back
         (Gram + the ridge walk's median + max |ncorr|);
      B  the conditioned null phenotypes (need M but not U), the null GEMM with its histogram epilogue,
         the observed histograms and one packed read-back (histograms + thresholds).
    The host meanwhile decomposes the n x n Gram (the only thing the device has to wait for), tests the
    observed phenotype, launches the PC regressions of all permutations and turns the histograms into
    the FDR table.  Phase B is queued before the decomposition when the permutation draw has already
    finished, after it otherwise."""
    out = select_output(show_progress)
    r, n = res.r, res.n
    dev = res.ncorr.device
    y = res.y_std
    ks = list(res.ks)
    ks_dev = np.unique(np.asarray(ks, dtype=np.int64))  # the device kernels want ascending, distinct ks
    ks_pos = np.searchsorted(ks_dev, np.asarray(ks, dtype=np.int64))
    kmax = int(ks_dev[-1])
    comm = res.comm
    Kl = min(1000, Nnull) if local_test else 0
    want_null_table = res.svd_top is None  # full result surface: every null p-value from scipy
    y_d = _to_dev(y)

    # ---- phase A ----
    def phase_a():
        G_d = _nam.gram_device(res.x, n, comm=comm, planes=res.planes)
        # the leading max(ks) eigenpairs of the Gram follow on the device (one launch); only the full result
        # surface (and n > 512) decomposes the Gram with LAPACK on the host
        eig = None
        if res.svd_top is not None and res.svd_top < n <= _nam.DEVICE_EIG_MAX_N and res.svd_top <= 64:
            # one CTA, ~1 ms: on a side stream, beside the null GEMM (which leaves it an SM, phase B)
            w_d = torch.empty(res.svd_top, dtype=torch.float64, device=dev)
            ut_d = torch.empty((res.svd_top, n), dtype=torch.float64, device=dev)
            gram_done = torch.cuda.Event()
            gram_done.record(_lib.current_stream_object())
            side = _side_stream(dev, "eig")
            with torch.cuda.stream(side):
                side.wait_event(gram_done)
                _lib.sym_eig_top(G_d, res.svd_top, w_d, ut_d)
                back_e = _Readback("eig", [w_d, ut_d])
            eig = (w_d, ut_d, back_e, G_d)  # (G_d stays referenced until the side stream is done with it)
        mx = torch.zeros(1, dtype=torch.float64, device=dev)
        tabs = None
        if local_test:
            _lib.absmax(res.ncorr, res.valid, mx)
            if comm is not None:
                comm.all_reduce(mx, op="max")
            thr_d = torch.empty(THRESHOLD_CAP, dtype=torch.float64, device=dev)
            edges_d = torch.empty(THRESHOLD_CAP, dtype=torch.float64, device=dev)
            count_d = torch.empty(1, dtype=torch.int32, device=dev)
            _lib.fdr_thresholds(mx, thr_d, edges_d, count_d)  # :101-102, _stats.py:51
            tabs = (thr_d, edges_d, count_d)
        med = res.ridge_median if res.ridge_median is not None else torch.zeros(2, dtype=torch.float64, device=dev)
        return _Readback("phase_a", [G_d, med, mx] if eig is None else [med, mx]), tabs, eig

    # ---- phase B ----
    def phase_b(tabs, eig_running=False):
        # permutations: indices from the host RNG (bit-exact), everything else on the device
        if perms is not None:
            perm_d = perms.result_device(dev)
        else:  # a shard other than rank 0: the indices are drawn once, by rank 0
            perm_d = torch.empty((Nnull, n), dtype=torch.int32, device=dev)
        if comm is not None and host_draw:
            comm.broadcast(perm_d, src=0)
        mark("permutations uploaded")
        C_d = _to_dev(res.C) if r else None
        W_d = _to_dev(np.ascontiguousarray(res.W_last)) if r else None
        back = None
        if local_test:  # :92-103
            print("computing neighborhood-level FDRs", file=out)
            thr_d, edges_d, count_d = tabs
            hist = torch.zeros(THRESHOLD_CAP, dtype=torch.int64, device=dev)  # summed over the Kl nulls
            obs = torch.zeros((2, THRESHOLD_CAP), dtype=torch.int32, device=dev)
            # ycond_ = M.y_[:, :Kl] / std (ddof=1) (:94-97) as transposed fp16 hi/lo planes, then
            # (cells x n) . (n x Kl) on the tensor cores with the histogram epilogue straight out of TMEM
            ytp = _lib.Planes(Kl, n, dev, zero=True)
            _lib.perm_stats(y_d, perm_d[:Kl], C_d, W_d, None, None, None, None, None, Kl, planes=ytp)
            cap = _lib.tc_max_ctas(_sm_count(dev) - 1) if eig_running else None  # an SM for the eigensolver
            try:
                _lib.null_hist_tc_dev(res.planes, n, ytp, Kl, edges_d, count_d, hist)
            finally:
                if cap is not None:
                    _lib.tc_max_ctas(cap)
            _lib.obs_hist_dev(res.ncorr, res.valid, edges_d, thr_d, count_d, obs[0], obs[1])
            if comm is not None:  # counts over all shards
                comm.all_reduce(hist)
                comm.all_reduce(obs)
            # the FDR table (_stats.py:79-80) and the per-cell FDR column follow on the device
            fdr_d = torch.empty(THRESHOLD_CAP, dtype=torch.float64, device=dev)
            pmin_d = torch.empty(THRESHOLD_CAP, dtype=torch.float64, device=dev)
            _lib.fdr_table(hist, obs[0], count_d, Kl, fdr_d, pmin_d)
            back = _Readback("phase_b", [hist, obs, thr_d, count_d, fdr_d])
            if columns is not None:
                columns.start_fdr(thr_d, pmin_d, count_d)
        mark("null kernels launched")
        return perm_d, C_d, W_d, back

    def launch_pc_regressions(U, perm_d, C_d, W_d, eig=None):
        """PC regressions of every permuted phenotype (:84) -> SSEs and, unless the full table of null
        p-values is wanted on the host, the F survival function + min over ks on the device.  ``eig``:
        the device-resident eigenpairs (U^T never left the GPU)."""
        ssered_d = torch.empty(Nnull, dtype=torch.float64, device=dev)
        ssefull_d = torch.full((Nnull, len(ks_dev)), float("nan"), dtype=torch.float64, device=dev)
        Ut_d = eig[1][:kmax] if eig is not None else _to_dev(np.ascontiguousarray(U[:, :kmax].T))
        ks_d = _to_dev(ks_dev.astype(np.int32))
        _lib.perm_stats(y_d, perm_d, C_d, W_d, Ut_d, ks_d, ssered_d, ssefull_d, None, 0)
        if want_null_table:
            return ssered_d, ssefull_d, None
        minp_d = torch.empty(Nnull, dtype=torch.float64, device=dev)
        argk_d = torch.empty(Nnull, dtype=torch.int32, device=dev)
        r2_d = torch.empty(Nnull, dtype=torch.float64, device=dev)
        _lib.perm_minp(ssered_d, ssefull_d, ks_d, n, r, minp_d, argk_d, r2_d)
        return ssered_d, ssefull_d, _Readback("minp", [minp_d, r2_d])

    def observed_test(U):
        """Observed phenotype (:64-74): n-sized host arithmetic in float64."""
        ycond = res.M.dot(y)
        ycond = ycond / ycond.std(ddof=1)  # a pandas Series in the reference -> ddof=1
        ssered = np.array([ycond.dot(ycond)])
        ssefull = np.array([[np.sum((U[:, :k].dot(U[:, :k].T.dot(ycond)) - ycond) ** 2) for k in ks]])
        p_all, r2_all = _f_pvalues(ssered, ssefull, ks, n, r)
        k, p, r2 = (a[0] for a in _pick(p_all, r2_all, ks))
        if k == max(ks):  # :65-67
            warnings.warn(("data supported use of {} NAM PCs, which is the maximum considered. "
                           'Consider allowing more PCs by using the "ks" argument.').format(k))
        beta = U[:, :k].T.dot(ycond)  # :72
        yhat = U[:, :k].dot(beta)
        r2_perpc = (beta / np.sqrt(ycond.dot(ycond))) ** 2  # :74
        return Namespace(k=k, p=p, r2=r2, beta=beta, yresid_hat=yhat, yresid=ycond, r2_perpc=r2_perpc)

    def global_pvalue(p, sse_d):
        """:84-88 from the regressions of the permuted phenotypes."""
        if sse_d[2] is None:  # every null p-value from scipy, in the caller's order of ks
            ssered, ssefull = sse_d[0].cpu().numpy(), sse_d[1].cpu().numpy()[:, ks_pos]
            nullp, nullr2 = _f_pvalues(ssered, ssefull, ks, n, r)
            _, nullminps, nullr2s = _pick(nullp, nullr2, ks)
        else:
            # min-p per permutation from the device (fp64 incomplete beta, ~1e-13 of scipy); the few that
            # fall within 1e-9 relative of the decision threshold are re-evaluated with scipy so that the
            # count below is exactly the reference's
            nullminps, nullr2s = (a.copy() for a in sse_d[2].get())
            thr = p + 1e-8
            near = np.abs(nullminps - thr) <= 1e-9 * thr
            if near.any():
                ix = torch.as_tensor(np.nonzero(near)[0], device=dev)
                pp, rr2 = _f_pvalues(sse_d[0][ix].cpu().numpy(), sse_d[1][ix].cpu().numpy()[:, ks_pos], ks, n, r)
                _, nullminps[near], nullr2s[near] = _pick(pp, rr2, ks)
        nhit = int((nullminps <= p + 1e-8).sum())
        if nhit == 0:
            warnings.warn("global association p-value attained minimal possible value. "
                          "Consider increasing Nnull")
        mark("global p done")
        return (nhit + 1) / (Nnull + 1), nullminps, nullr2s

    def fdr_table(back):
        """:105-118 from the histograms (and the thresholds the device derived from max |ncorr|)."""
        hist_h, obs_h, thr_h, count_h, fdr_h = back.get()
        T = int(count_h[0])
        if T >= THRESHOLD_CAP:
            raise _lib.CnaError(f"more than {THRESHOLD_CAP - 1} FDR thresholds")
        thresholds = thr_h[:T].copy()
        fdr_vals = fdr_h[:T].copy()  # _stats.py:64-83 (cna_fdr_table; == _stats.fdr_from_counts(hist, obs[0], Kl))
        num_detected = _stats.tails_from_hist(obs_h[1, :T].astype(np.int64))  # :105-108
        # the DataFrame is part of the full result surface only
        fdrs = (pd.DataFrame({"threshold": thresholds, "fdr": fdr_vals, "num_detected": num_detected})
                if want_null_table else True)
        t5 = t10 = None
        if not np.nanmin(fdr_vals) > 0.05:  # :111-114 (Series.min skips NaN; first row with fdr <= 0.05)
            t5 = thresholds[np.nonzero(fdr_vals <= 0.05)[0][0]]
        if not np.nanmin(fdr_vals) > 0.1:  # :115-118
            t10 = thresholds[np.nonzero(fdr_vals <= 0.1)[0][0]]
        mark("fdr table done")
        return fdrs, t5, t10

    back_a, tabs, eig = phase_a()
    if columns is not None:
        columns.start_coef()
    # phase B needs the permutation indices: queued now when the draw has already finished (or when this
    # rank only receives them), otherwise as soon as the Gram is here if the draw has finished by then,
    # otherwise after the decomposition
    # (with the eigensolver on the device the host has nothing to do for the Gram: it waits for the draw
    # right here and queues phase B behind phase A without a gap)
    b = phase_b(tabs, eig is not None) if (perms is None or perms.done() or eig is not None) else None
    while True:
        got = back_a.get()
        mark("gram on host" if eig is None else "ridge median on host")
        if res.settle(float(got[-2][0])):
            break
        # the first ridge did not bring the median batch kurtosis down to 6 (_nam.py:154): the walk has
        # now been finished ridge by ridge and everything queued on the speculative NAM is repeated
        back_a, tabs, eig = phase_a()
        if columns is not None:
            columns.start_coef()
        b = phase_b(tabs, eig is not None) if b is not None else None
    if b is None and perms.done():
        b = phase_b(tabs, eig is not None)
    if eig is None:
        U, svs, res.G = _nam.svd_of_gram(got[0].copy(), res.svd_top)
        if b is None:
            b = phase_b(tabs)
        perm_d, C_d, W_d, back_b = b
        sse_d = launch_pc_regressions(U, perm_d, C_d, W_d)  # the device works on these while the host ...
    else:
        if b is None:
            b = phase_b(tabs, True)
        perm_d, C_d, W_d, back_b = b
        # the PC regressions are queued behind the null GEMM and wait (on the device) for the eigenvectors
        _lib.current_stream_object().wait_event(eig[2].event)
        sse_d = launch_pc_regressions(None, perm_d, C_d, W_d, eig)
        w_h, ut_h = eig[2].get()
        mark("eigenpairs on host")
        # same shapes as svd_of_gram: trailing columns / values are zero and are never read
        U, svs = np.zeros((n, n)), np.zeros(n)
        U[:, :res.svd_top] = ut_h.T
        svs[:res.svd_top] = np.maximum(w_h, 0.0)
    res.U, res.svs = U, svs
    o = observed_test(U)                                 # ... tests the observed phenotype
    fdrs, fdr_5p_t, fdr_10p_t = fdr_table(back_b) if local_test else (None, None, None)
    pfinal, nullminps, nullr2s = global_pvalue(o.p, sse_d)

    return Namespace(p=pfinal, nullminps=nullminps, k=o.k, ncorrs=None, fdrs=fdrs,
                     fdr_5p_t=fdr_5p_t, fdr_10p_t=fdr_10p_t, yresid_hat=o.yresid_hat, yresid=o.yresid,
                     ks=res.ks, beta=o.beta, r2=o.r2, r2_perpc=o.r2_perpc,
                     nullr2_mean=nullr2s.mean(), nullr2_std=nullr2s.std())


def association(data, y, sid_name, batches=None, covs=None, donorids=None, ks=None, key_added="coef",
                max_frac_pcs=0.15, nsteps=None, show_progress=False, allow_low_sample_size=False,
                return_full=False, ridges=None, **kwargs):
    """``_association.py:193-242``.  Returns the global p-value, or the full result Namespace when
    ``return_full``; writes ``data.obs[key_added]`` and ``data.obs[key_added + '_fdr']``.

    Limits (checked before any device work): at most 1024 samples in ``data.obs[sid_name]`` and at
    most 1024 selected samples."""
    out = select_output(show_progress)
    bad = set(kwargs) - {"Nnull", "force_permute_all", "local_test", "seed", "_host_draw"}
    if bad:  # the reference forwards **kwargs to _association(), which rejects anything else
        raise TypeError(f"_association() got an unexpected keyword argument '{sorted(bad)[0]}'")
    mark("association() entered")
    for name, val, kind in (("y", y, pd.Series), ("batches", batches, pd.Series), ("covs", covs, pd.DataFrame),
                            ("donorids", donorids, pd.Series)):
        if val is not None and not isinstance(val, kind):  # :132-139, before any device work
            raise TypeError(f"'{name}' must be a pandas {kind.__name__}, but got {type(val)}")
    # one factorisation of the sample-id column serves the input checks (:140-143) and the NAM (:51)
    codes = _graph.sample_codes(data, sid_name)
    if len(codes[0]) > 1024:
        raise ValueError(f"cna_b200 supports at most 1024 samples (data.obs['{sid_name}'] has {len(codes[0])})")
    # The diffusion needs nothing but the graph and the sample codes.  With the graph resident it is queued
    # first, and the input checks, the design algebra and the permutation draw run on the host while the
    # device works (an exception below simply abandons the queued kernels: nothing has been written to
    # data.obs).  A host graph is uploaded and reordered first (~10 ms): there the draw is started before.
    user_batches = batches

    # Every sample of the data selected, in label order (the common call): the QC statistic of a cell is then
    # a by-product of the residualisation pass (its batch-mean functionals are the QC's batch means), so the
    # last diffusion step runs without its QC epilogue and the decision is applied afterwards (cna_qc_fixup)
    qc_in_pass = (not show_progress and not os.environ.get("CNA_B200_QC_IN_SPMM")
                  and _all_samples_selected(codes[0], y, batches, covs))

    def launch_nam():
        print("computing NAM", file=out)
        st = _nam._nam_device(data, sid_name, nsteps=nsteps, show_progress=show_progress, codes=codes,
                              qc_batches=None if qc_in_pass else user_batches)
        mark("diffusion launched")
        return st

    resident = isinstance(getattr(data, "graph", None), _graph.DeviceGraph)
    stn = launch_nam() if resident else None
    batches, filter_samples = check_inputs(data, y, sid_name, batches, covs, donorids,
                                           allow_low_sample_size, present=codes[0])
    mark("check_inputs done")
    Nnull = kwargs.get("Nnull", 1000)
    local_test = kwargs.get("local_test", True)

    # ---- sample bookkeeping (:178-191) and the small design algebra, all on the host ----
    fs = np.asarray(filter_samples, dtype=bool)
    sids = y.index[fs]
    n = int(fs.sum())
    batches_f = batches.reindex(y.index).to_numpy()[fs]
    covs_f = covs.reindex(y.index).to_numpy()[fs] if covs is not None else None
    donor_f = donorids.reindex(y.index).to_numpy()[fs] if donorids is not None else None
    y_f = np.asarray(y.to_numpy()[fs], dtype=np.float64)
    y_std = (y_f - y_f.mean()) / y_f.std()  # :22 (ndarray -> ddof=0)
    npcs = min(n, max([10] + [int(max_frac_pcs * n)] + [ks if ks is not None else []][0]))  # :207
    ks_eff = default_ks(n) if ks is None else ks
    design = _nam.design_matrix(covs_f, batches_f, n)
    r = design[0].shape[1]

    if kwargs.get("seed") is not None:
        np.random.seed(kwargs["seed"])  # :15-16
    if max(ks_eff) + r >= n:  # :29-33 (the reference raises this after seeding, before any draw)
        raise ValueError(
            "Maximum number of PCs plus number of covariates must be less than n-1. "
            f"Currently it is {max(ks_eff) + r} while n is {n}. Either reduce the number of covariates "
            "or reduce the number of PCs to consider using the optional argument ks=[...].")
    perm_batches = np.ones(n) if kwargs.get("force_permute_all", False) else batches_f  # :17-18
    comm = getattr(data, "comm", None)
    if comm is not None and return_full:
        raise NotImplementedError("return_full=True is not supported on a cell-axis shard")
    # the permutation draw only needs the sample-level inputs.  Device engine (default): queued on a side
    # stream, every rank of a sharded run draws the same matrix itself; host engine: rank 0 draws on the
    # host's threads and broadcasts
    dev = _graph.device()
    host_draw = kwargs.get("_host_draw", False) or not _stats._device_draw_enabled()
    perms = (_stats.PermutationDraw(y_std, perm_batches, donor_f, Nnull, device=dev,
                                    engine="host" if host_draw else "device")
             if (not host_draw or comm is None or comm.rank == 0) else None)
    mark("permutation draw started")
    try:
        if stn is None:
            stn = launch_nam()
        # ---- QC and residualisation: queued behind the diffusion, no host round trip ----
        if qc_in_pass:
            stn._keep, stn.qc_median, stn.qc_kurt = None, None, None
        else:
            _nam._qc_device(stn, batches, show_progress=show_progress)
        colmap = stn.labels.get_indexer(sids)  # NAM.reindex(y.index)[filter_samples], :178-181
        res = _nam.resid_nam_device(stn, colmap, covs_f, batches_f, y_std, ridges=ridges,
                                    show_progress=show_progress, want_x=return_full, speculate=True,
                                    design=design, qc_in_pass=qc_in_pass)
        mark("resid pass launched")
        res.y_std = y_std
        res.ks = ks_eff
        # only the leading max(ks) components are read unless the full result surface is requested
        res.svd_top = None if return_full else int(max(ks_eff))
        print("performing association test", file=out)
        # every rank of a sharded run ends with the full columns in its copy of data.obs
        columns = _Columns(data, key_added, stn, res, gather=True)
        core = _association(res, perms, Nnull=Nnull, local_test=local_test, show_progress=show_progress,
                            columns=columns, host_draw=host_draw)
    except BaseException:
        if perms is not None:
            try:
                perms.cancel()  # joins the draw and hands numpy's generator its advanced state back
            except _stats.RedoWithHostDraw:
                pass
        raise
    if perms is not None:
        try:
            perms.cancel()
        except _stats.RedoWithHostDraw:
            # the device draw could not certify one of its argsorts (~1e-6 per call): nothing has been written
            # to data.obs yet and the generator is back in its state from before the draw — once more with
            # the host engine
            return association(data, y, sid_name, batches=user_batches, covs=covs, donorids=donorids, ks=ks,
                               key_added=key_added, max_frac_pcs=max_frac_pcs, nsteps=nsteps,
                               show_progress=show_progress, allow_low_sample_size=allow_low_sample_size,
                               return_full=return_full, ridges=ridges,
                               **{**{k: v for k, v in kwargs.items() if k != "seed"}, "_host_draw": True})
    columns.write()
    svs = res.svs
    LAST.__dict__.clear()
    LAST.__dict__.update(svs=np.array(svs[:min(8, len(svs))]), k=core.k, fdr_5p_t=core.fdr_5p_t,
                         fdr_10p_t=core.fdr_10p_t, p=core.p, n=n, r=res.r, ridge_log=list(res.ridge_log))
    if not return_full:
        report()
        return core.p

    # ---- full result surface (_nam.py:168-175, _association.py:223-225) ----
    dev = res.ncorr.device
    vmask = _nam.to_caller_order(stn, res.valid).bool()
    kept = vmask.cpu().numpy()
    cells = data.obs.index[kept]
    pcs = ["PC" + str(i) for i in range(1, n + 1)]
    full = Namespace()
    full.M = pd.DataFrame(res.M, index=sids, columns=sids)
    full.r = res.r
    xk = _nam.to_caller_order(stn, res.x)[vmask][:, :n]
    full.namresid = pd.DataFrame(xk.t().double().cpu().numpy(), index=sids, columns=cells)
    full.namresid_sampleXpc = pd.DataFrame(res.U, index=sids, columns=pcs)
    V = _nam.nbhd_loadings(res.x, n, res.U, svs, planes=res.planes,
                           rows=lambda v: _nam.to_caller_order(stn, v)[vmask])
    full.namresid_nbhdXpc = pd.DataFrame(V, index=cells, columns=pcs)
    full.namresid_svs = pd.Series(svs, index=pcs)[:npcs]
    full.namresid_varexp = pd.Series(svs / n / len(cells), index=pcs)
    full.__dict__.update(vars(core))
    full.ncorrs = pd.Series(data.obs[key_added].to_numpy()[kept], index=cells)
    full.yresid = pd.Series(core.yresid, index=sids)
    cm = torch.as_tensor(colmap, device=dev, dtype=torch.long)
    nam_sel = (_nam.to_caller_order(stn, stn.s)[vmask][:, cm].double() * stn.inv_count[cm])
    full.nam = pd.DataFrame(nam_sel.t().cpu().numpy(), index=sids, columns=cells)
    full.kept = kept
    return full
