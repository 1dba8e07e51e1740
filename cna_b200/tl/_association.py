"""Association driver: global permutation test + neighbourhood-level FDRs.

Mirrors ``src/cna/tools/_association.py`` of the reference: same signature, side effects on
``data.obs``, exceptions, warnings and result Namespace.  The NAM never leaves the GPU unless
``return_full=True`` asks for the big matrices.
"""
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


def _to_host_pinned(t, slot="results"):
    """Device tensor -> numpy array through a cached page-locked staging buffer (a pageable D2H of
    the two per-cell float64 columns costs more than the kernels that produced them).  The returned
    array is a view of the staging buffer of that ``slot``: consume it before the next call."""
    key = (t.dtype, tuple(t.shape))
    held = _PINNED.get(slot)
    buf = held[1] if held is not None and held[0] == key else None
    if buf is None:
        buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        _PINNED[slot] = (key, buf)
    buf.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return buf.numpy()


def _to_host_owned(t):
    """Device tensor -> numpy array in a page-locked buffer of its own (torch's pinned-memory cache
    hands the block of the column written by the previous call back once pandas lets go of it)."""
    buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    buf.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return buf.numpy()


def _adopt_column(obs, key, values):
    """``obs[key] = values`` without copying the N values a second time: the column adopts
    ``values`` (an array nothing else writes to; the ndarray keeps its buffer alive), where assigning
    the bare ndarray makes pandas copy it (~1 ms per million cells, on the critical path of a call)."""
    try:
        obs[key] = pd.Series(values, index=obs.index, copy=False)
    except Exception:  # noqa: BLE001 - an obs container that cannot adopt a Series gets the plain copy
        obs[key] = values


def default_ks(n):
    """``_association.py:25-28``."""
    incr = max(int(0.02 * n), 1)
    maxnpcs = max(min(4 * incr, int(n / 5)), 1)
    return np.arange(incr, maxnpcs + 1, incr)


_POOL = None
_HELPER = None


def _helper_pool():
    global _HELPER
    if _HELPER is None:
        from concurrent.futures import ThreadPoolExecutor
        _HELPER = ThreadPoolExecutor(max_workers=1, thread_name_prefix="cna-svd")
    return _HELPER



def _f_sf(f, dfn, dfd):
    """``st.f.sf`` (:46).  scipy's incomplete-beta ufunc releases the GIL, so large inputs are
    evaluated in row chunks on a small thread pool (same function, same bits)."""
    global _POOL
    if f.shape[0] < 2048:
        return st.f.sf(f, dfn, dfd)
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


def _association(res, perms, Nnull=1000, local_test=True, show_progress=False, idle_work=None):
    """``_nam.py:163`` (Gram + SVD of the residualised NAM) and ``_association.py:10-129`` after
    seeding / permutation drawing (done by the caller so that they overlap with the NAM kernels).
    ``res`` carries the device-resident residualised NAM (``res.planes`` / ``res.x``), M, r, the
    standardised phenotype and ks; U, svs and the Gram are added to it.

    Order of work: the Gram and max|ncorr| are launched and read back with one sync; the n x n SVD then
    runs on a helper thread.  Meanwhile this thread launches the conditioned null phenotypes (which need
    M but not U) and the null GEMM + histograms, turns the histograms into the FDR table and starts on
    ``idle_work(fdrs, svd_done)`` (the per-cell output columns; a sequence of steps that stops early once
    ``svd_done()`` is true).  With U: observed statistics, the PC regressions of all permutations are
    launched, ``idle_work(fdrs, None)`` finishes its remaining steps while they run, then the global
    p-value."""
    out = select_output(show_progress)
    M, r, n = res.M, res.r, res.n
    dev = res.ncorr.device
    y = res.y_std
    ks = res.ks
    kmax = int(max(ks))
    comm = res.comm
    Kl = min(1000, Nnull) if local_test else 0

    G_d = _nam.gram_device(res.x, n, comm=comm, planes=res.planes)
    mx = torch.zeros(1, dtype=torch.float64, device=dev)
    if local_test:
        _lib.absmax(res.ncorr, res.valid, mx)
        if comm is not None:
            comm.all_reduce(mx, op="max")
    Gh = G_d.cpu().numpy()
    mark("gram on host")
    want_null_table = res.svd_top is None  # full result surface: every null p-value from scipy

    def launch_pc_regressions(U):
        """PC regressions of every permuted phenotype (:84) -> SSEs and, unless the full table of null
        p-values is wanted on the host, the F survival function + min over ks on the device."""
        ssered_d = torch.empty(Nnull, dtype=torch.float64, device=dev)
        ssefull_d = torch.empty((Nnull, len(ks)), dtype=torch.float64, device=dev)
        Ut_d = _to_dev(np.ascontiguousarray(U[:, :kmax].T))
        ks_d = _to_dev(np.asarray(ks, dtype=np.int32))
        _lib.perm_stats(y_d, perm_d, C_d, W_d, Ut_d, ks_d, ssered_d, ssefull_d, None, 0)
        if want_null_table:
            return ssered_d, ssefull_d, None
        minp_d = torch.empty(Nnull, dtype=torch.float64, device=dev)
        argk_d = torch.empty(Nnull, dtype=torch.int32, device=dev)
        r2_d = torch.empty(Nnull, dtype=torch.float64, device=dev)
        _lib.perm_minp(ssered_d, ssefull_d, ks_d, n, r, minp_d, argk_d, r2_d)
        return ssered_d, ssefull_d, (minp_d, r2_d)

    def fetch(sse_d):
        if sse_d[2] is None:
            return (sse_d[0].cpu().numpy(), sse_d[1].cpu().numpy(), None)
        return (sse_d[0], sse_d[1], (sse_d[2][0].cpu().numpy(), sse_d[2][1].cpu().numpy()))

    def observed_test(U):
        """Observed phenotype (:64-74): n-sized host arithmetic in float64."""
        ycond = M.dot(y)
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
        sse_host = fetch(sse_d)
        if sse_host[2] is None:
            nullp, nullr2 = _f_pvalues(sse_host[0], sse_host[1], ks, n, r)
            _, nullminps, nullr2s = _pick(nullp, nullr2, ks)
        else:
            # min-p per permutation from the device (fp64 incomplete beta, ~1e-13 of scipy); the few that
            # fall within 1e-9 relative of the decision threshold are re-evaluated with scipy so that the
            # count below is exactly the reference's
            nullminps, nullr2s = sse_host[2]
            thr = p + 1e-8
            near = np.abs(nullminps - thr) <= 1e-9 * thr
            if near.any():
                ix = torch.as_tensor(np.nonzero(near)[0], device=dev)
                pp, rr2 = _f_pvalues(sse_host[0][ix].cpu().numpy(), sse_host[1][ix].cpu().numpy(), ks, n, r)
                _, nullminps[near], nullr2s[near] = _pick(pp, rr2, ks)
        nhit = int((nullminps <= p + 1e-8).sum())
        if nhit == 0:
            warnings.warn("global association p-value attained minimal possible value. "
                          "Consider increasing Nnull")
        mark("global p done")
        return (nhit + 1) / (Nnull + 1), nullminps, nullr2s

    # The n x n SVD (_nam.py:105) needs only the Gram: it runs on a helper thread (LAPACK releases the
    # GIL) while this thread launches the null kernels and writes the per-cell outputs.
    svd_future = _helper_pool().submit(_nam.svd_of_gram, Gh, res.svd_top)

    fdrs, fdr_5p_t, fdr_10p_t = None, None, None
    try:
        # ---- permutations: indices from the host RNG (bit-exact), everything else on the device ----
        if perms is not None:
            perm_d = perms.result_device(dev)
        else:  # a shard other than rank 0: the indices are drawn once, by rank 0
            perm_d = torch.empty((Nnull, n), dtype=torch.int32, device=dev)
        if comm is not None:
            comm.broadcast(perm_d, src=0)
        mark("permutations uploaded")
        y_d = _to_dev(y)
        C_d = _to_dev(res.C) if r else None
        W_d = _to_dev(np.ascontiguousarray(res.W_last)) if r else None

        # ---- neighbourhood-level null (:92-103) ----
        if local_test:
            print("computing neighborhood-level FDRs", file=out)
            maxcorr = max(float(mx.item()), 0.001)  # :101
            thresholds = np.arange(maxcorr / 4, maxcorr, maxcorr / 400)  # :102
            edges = _stats.threshold_edges(thresholds)
            T = len(thresholds)
            edges_d, thr_d = _to_dev(edges), _to_dev(thresholds)
            hist = torch.zeros(T, dtype=torch.int64, device=dev)  # summed over the Kl nulls
            obs = torch.zeros((2, T), dtype=torch.int32, device=dev)
            # ycond_ = M.y_[:, :Kl] / std (ddof=1) (:94-97) as transposed fp16 hi/lo planes, then
            # (cells x n) . (n x Kl) on the tensor cores with the histogram epilogue straight out of TMEM
            ytp = _lib.Planes(Kl, n, dev, zero=True)
            _lib.perm_stats(y_d, perm_d[:Kl], C_d, W_d, None, None, None, None, None, Kl, planes=ytp)
            _lib.null_hist_tc(res.planes, n, ytp, Kl, edges_d, float(edges[0]), hist)
            _lib.obs_hist(res.ncorr, res.valid, edges_d, thr_d, obs[0], obs[1])
            if comm is not None:  # counts over all shards
                comm.all_reduce(hist)
                comm.all_reduce(obs)
        mark("null kernels launched")

        if local_test:  # :105-118; needs the histograms only
            obs_h = obs.cpu().numpy()
            fdr_vals = _stats.fdr_from_counts(hist.cpu().numpy(), obs_h[0], n_null=Kl)  # _stats.py:64-83
            num_detected = _stats.tails_from_hist(obs_h[1].astype(np.int64))  # :105-108
            fdrs = pd.DataFrame({"threshold": thresholds, "fdr": fdr_vals, "num_detected": num_detected})
            if not np.nanmin(fdr_vals) > 0.05:  # :111-114 (Series.min skips NaN; first row with fdr <= 0.05)
                fdr_5p_t = thresholds[np.nonzero(fdr_vals <= 0.05)[0][0]]
            if not np.nanmin(fdr_vals) > 0.1:  # :115-118
                fdr_10p_t = thresholds[np.nonzero(fdr_vals <= 0.1)[0][0]]
            mark("fdr table done")
        if idle_work is not None:
            idle_work(fdrs, svd_future.done)  # per-cell outputs, for as long as the SVD thread is busy
    except BaseException:
        svd_future.result()
        raise
    U, svs, res.G = svd_future.result()
    res.U, res.svs = U, svs
    o = observed_test(U)
    sse_d = launch_pc_regressions(U)
    if idle_work is not None:
        idle_work(fdrs, None)  # whatever is left of them, while the GPU runs the regressions
    pfinal, nullminps, nullr2s = global_pvalue(o.p, sse_d)

    return Namespace(p=pfinal, nullminps=nullminps, k=o.k, ncorrs=None, fdrs=fdrs,
                     fdr_5p_t=fdr_5p_t, fdr_10p_t=fdr_10p_t, yresid_hat=o.yresid_hat, yresid=o.yresid,
                     ks=ks, beta=o.beta, r2=o.r2, r2_perpc=o.r2_perpc,
                     nullr2_mean=nullr2s.mean(), nullr2_std=nullr2s.std())


def association(data, y, sid_name, batches=None, covs=None, donorids=None, ks=None, key_added="coef",
                max_frac_pcs=0.15, nsteps=None, show_progress=False, allow_low_sample_size=False,
                return_full=False, ridges=None, **kwargs):
    """``_association.py:193-242``.  Returns the global p-value, or the full result Namespace when
    ``return_full``; writes ``data.obs[key_added]`` and ``data.obs[key_added + '_fdr']``."""
    out = select_output(show_progress)
    bad = set(kwargs) - {"Nnull", "force_permute_all", "local_test", "seed"}
    if bad:  # the reference forwards **kwargs to _association(), which rejects anything else
        raise TypeError(f"_association() got an unexpected keyword argument '{sorted(bad)[0]}'")
    mark("association() entered")
    for name, val, kind in (("y", y, pd.Series), ("batches", batches, pd.Series), ("covs", covs, pd.DataFrame),
                            ("donorids", donorids, pd.Series)):
        if val is not None and not isinstance(val, kind):  # :132-139, before any device work
            raise TypeError(f"'{name}' must be a pandas {kind.__name__}, but got {type(val)}")
    # one factorisation of the sample-id column serves the input checks (:140-143) and the NAM (:51)
    codes = _graph.sample_codes(data, sid_name)
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
    r = _nam.design_matrix(covs_f, batches_f, n)[0].shape[1]

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
    # the permutation draws only need the sample-level inputs: start them before any GPU work
    perms = (_stats.PermutationDraw(y_std, perm_batches, donor_f, Nnull)
             if comm is None or comm.rank == 0 else None)
    mark("permutation draw started")

    # ---- launch the diffusion (asynchronous unless nsteps is None) ----
    print("computing NAM", file=out)
    try:
        stn = _nam._nam_device(data, sid_name, nsteps=nsteps, show_progress=show_progress, codes=codes,
                               qc_batches=batches)
    except BaseException:
        if perms is not None:
            perms.cancel()
        raise
    mark("diffusion launched")

    # ---- QC, residualisation, Gram + SVD ----
    _nam._qc_device(stn, batches, show_progress=show_progress)
    mark("QC done (first sync)")
    colmap = stn.labels.get_indexer(sids)  # NAM.reindex(y.index)[filter_samples], :178-181
    res = _nam.resid_nam_device(stn, colmap, covs_f, batches_f, y_std, ridges=ridges,
                                show_progress=show_progress, want_x=return_full)
    mark("resid pass done")
    res.y_std = y_std
    res.ks = ks_eff
    # only the leading max(ks) components are read unless the full result surface is requested
    res.svd_top = None if return_full else int(max(ks_eff))
    print("performing association test", file=out)
    N = stn.N
    dev = res.ncorr.device
    columns_written = []

    def fdr_lookup_tables(fdrs):
        if fdrs is None:
            # local_test=False: the reference writes the coefficients and then crashes looking up FDRs
            # (res.fdrs is None at :235); here the FDR column is simply not written.
            return np.array([np.inf]), np.array([1.0])
        return fdrs.threshold.to_numpy(), np.fmin.accumulate(fdrs.fdr.to_numpy())  # Series.min() skips NaN

    def column_steps(fdrs):
        """data.obs[key_added] (:228-231) is known as soon as the NAM is residualised and the per-cell
        FDR (:234-237) as soon as the null histograms are: both are copied back and written in the
        shadow of the SVD thread and of the permutation regressions.  One yield per step."""
        coef = torch.where(res.valid.bool(), res.ncorr, torch.full_like(res.ncorr, float("nan")))
        if stn.graph is not None:
            coef = stn.graph.unpermute(coef)
        host = _to_host_owned(coef)
        yield
        if key_added in data.obs:
            warnings.warn(f"Key '{key_added}' already exists in data.obs. Overwriting.")
        _adopt_column(data.obs, key_added, host)
        mark("coef column written")
        yield
        if fdrs is not None:
            thr, pmin = fdr_lookup_tables(fdrs)
            coef_d = torch.empty(N, dtype=torch.float64, device=dev)
            fdr_d = torch.empty(N, dtype=torch.float64, device=dev)
            _lib.cell_fdr(res.ncorr, res.valid, _to_dev(thr), _to_dev(pmin), coef_d, fdr_d)
            fdr_h = _to_host_owned(fdr_d if stn.graph is None else stn.graph.unpermute(fdr_d))
            yield
            _adopt_column(data.obs, f"{key_added}_fdr", fdr_h)
        columns_written.append(True)
        mark("obs written")

    def write_columns(fdrs, stop):
        """Runs the remaining steps, returning early (to be resumed) once ``stop()`` is true."""
        if not steps:
            steps.append(column_steps(fdrs))
        for _ in steps[0]:
            if stop is not None and stop():
                return

    steps = []
    early = comm is None and not return_full
    core = _association(res, perms, Nnull=Nnull, local_test=local_test, show_progress=show_progress,
                        idle_work=write_columns if early else None)
    svs = res.svs
    if columns_written:
        report()
        return core.p

    # ---- neighbourhood-level outputs (:228-237) ----
    coef_d = torch.empty(N, dtype=torch.float64, device=dev)
    fdr_d = torch.empty(N, dtype=torch.float64, device=dev)
    if key_added in data.obs:
        warnings.warn(f"Key '{key_added}' already exists in data.obs. Overwriting.")
    thr, pmin = fdr_lookup_tables(core.fdrs)
    _lib.cell_fdr(res.ncorr, res.valid, _to_dev(thr), _to_dev(pmin), coef_d, fdr_d)
    both = torch.stack([coef_d, fdr_d], dim=1)  # [cells x 2]
    if comm is not None:  # every rank ends with the full per-cell columns
        pad = torch.zeros((stn.rows_per, 2), dtype=torch.float64, device=dev)
        pad[:N] = both
        both = comm.all_gather_rows(pad)[: len(data.obs)]
    if stn.graph is not None:
        both = stn.graph.unpermute(both)  # back to the caller's cell order
    both = _to_host_pinned(both.t())
    mark("results on host")
    data.obs[key_added] = both[0]  # pandas copies on assignment (`both` is the reusable staging buffer)
    if core.fdrs is not None:
        data.obs[f"{key_added}_fdr"] = both[1]
    if not return_full:
        mark("obs written")
        report()
        return core.p

    # ---- full result surface (_nam.py:168-175, _association.py:223-225) ----
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
    full.ncorrs = pd.Series(both[0][kept], index=cells)
    full.yresid = pd.Series(core.yresid, index=sids)
    cm = torch.as_tensor(colmap, device=dev, dtype=torch.long)
    nam_sel = (_nam.to_caller_order(stn, stn.s)[vmask][:, cm].double() * stn.inv_count[cm])
    full.nam = pd.DataFrame(nam_sel.t().cpu().numpy(), index=sids, columns=cells)
    full.kept = kept
    return full
