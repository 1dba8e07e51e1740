"""TEST INFRASTRUCTURE — CPU oracle for the CNA NAM + permutation-association hot path.

A numpy/scipy/pandas restatement (float64 end to end) of what ``immunogenomics/cna`` v0.2.3 computes
in ``cna.tl.nam`` / ``cna.tl.association``.  Every function cites the reference ``file:line`` it
follows (paths relative to the reference checkout, ``src/cna/tools/``).  It is written in the
cells x samples (N x S) layout that the CUDA kernels use, i.e. the transpose of the reference's
DataFrames, so every per-neighbourhood operation is row-local.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module, and only as the checker or the timed CPU arm — never as a product code
path.  ``cna_b200`` itself must not import it.

Parity pinning: the reference ships no tests.  The oracle is pinned (tests/test_oracle_golden.py)
against (a) the printed outputs of the reference's ``demo/demo.ipynb`` on ``demo/data.h5ad``
(p = 0.000999000999000999, 9555 neighbourhoods at FDR 5 %), and (b) outputs of the unmodified
reference run in the build container through ``oracle/ref_shim.py`` and committed under
``tests/golden/`` by ``tests/golden/make_golden.py``.

Third-party arithmetic on the path (not vendored in the reference; minimum pins only in its
``setup.cfg:22-29``): scipy ``csr_matvecs`` / ``stats.kurtosis`` / ``special.fdtrc``, numpy legacy
``RandomState`` + ``argsort`` + ``histogram``, LAPACK ``dgesdd``/``dgesv``.  The oracle calls the
very same library entry points, so "reference source + installed numpy/scipy" is the operational
definition of correct.

Two execution modes:
  * default (vectorised): used by the parity tests — same numbers, seconds instead of minutes;
  * ``faithful=True``: keeps the reference's cost structure (per-step kurtosis and R^2 diagnostics,
    dense ``M`` products, Python loop over permutations, per-null ``np.histogram``, per-cell FDR
    lookup through ``Series.apply``) so that it can stand in for the reference as the timed CPU arm.
"""
import warnings
from argparse import Namespace

import numpy as np
import pandas as pd
import scipy.sparse as sp
import scipy.stats as st

DEFAULT_RIDGES = [1e5, 1e4, 1e3, 1e2, 1e1, 1e0, 1e-1, 1e-2, 1e-3, 1e-4, 0]


# --------------------------------------------------------------------------------------------
# graph / sample bookkeeping
# --------------------------------------------------------------------------------------------
def get_connectivity(data):
    """_nam.py:12-19 — modern AnnData keeps the kNN graph in ``.obsp``; legacy in ``.uns``."""
    obsp = getattr(data, "obsp", None)
    if obsp is not None and "connectivities" in obsp:
        return obsp["connectivities"]
    return data.uns["neighbors"]["connectivities"]


def sample_codes(sid):
    """Column order of ``pd.get_dummies(data.obs[sid_name])`` (_nam.py:51): categories for a
    categorical column, otherwise the sorted unique values.  Returns (labels, int codes)."""
    if isinstance(sid.dtype, pd.CategoricalDtype):
        labels = sid.cat.categories
        codes = sid.cat.codes.to_numpy()
    else:
        codes, labels = pd.factorize(sid, sort=True)
    return pd.Index(labels), np.asarray(codes, dtype=np.int64)


# --------------------------------------------------------------------------------------------
# diffusion (kernel i)
# --------------------------------------------------------------------------------------------
def diffuse_stepwise(A, s, maxnsteps=15, self_weight=1):
    """_nam.py:21-34.  s <- A.(s/colsums) + w.s/colsums, colsums = column sums of A + w."""
    colsums = np.asarray(A.sum(axis=0)).ravel() + self_weight  # :28
    s = np.asarray(s, dtype=np.float64)
    for _ in range(maxnsteps):
        scaled = s / colsums[:, None]
        s = A.dot(scaled) + self_weight * scaled  # :33
        yield s


def diffuse(A, s, nsteps, self_weight=1):
    """_nam.py:36-41."""
    for s in diffuse_stepwise(A, s, maxnsteps=nsteps, self_weight=self_weight):
        pass
    return s


def row_kurtosis_median(x):
    """_nam.py:59 — median over cells of the (biased, Fisher) kurtosis across samples."""
    return np.median(st.kurtosis(x, axis=1))


def _pearson_r2_cols(a, b):
    """_nam.py:47-49,60 — per-sample squared Pearson r between consecutive states (print-only)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        r = ((a - a.mean(axis=0)) * (b - b.mean(axis=0))).mean(axis=0) / a.std(axis=0) / b.std(axis=0)
    return r ** 2


def nam_raw(A, codes, nsamples, nsteps=None, maxnsteps=15, self_weight=1, diagnostics=None,
            faithful=False):
    """_nam.py:44-76.  Returns the N x S matrix ``s / C`` (transpose of the reference's ``snorm``).

    ``diagnostics`` (a dict) receives ``medkurt`` (the +3 values the reference prints at :62),
    ``r2_p20`` (:63, only when ``faithful``) and ``nsteps``.
    """
    N = len(codes)
    onehot = np.zeros((N, nsamples))
    onehot[np.arange(N), codes] = 1.0  # :51
    C = onehot.sum(axis=0)  # :54
    medkurts, r2s = [], []
    prev = np.inf
    old = np.zeros_like(onehot)
    need_stats = nsteps is None or faithful or diagnostics is not None
    s = onehot
    for i, s in enumerate(diffuse_stepwise(A, onehot, maxnsteps=maxnsteps, self_weight=self_weight)):
        if need_stats:
            with np.errstate(divide="ignore", invalid="ignore"):
                medkurt = row_kurtosis_median(s / C)  # :59
            medkurts.append(medkurt + 3)
            if faithful:
                r2s.append(np.percentile(_pearson_r2_cols(s, old), 20))  # :60,63
                old = s
        if nsteps is None:
            if prev - medkurt < 3 and i + 1 >= 3:  # :65
                break
            prev = medkurt
        elif i + 1 == nsteps:  # :69
            break
    if diagnostics is not None:
        diagnostics.update(medkurt=medkurts, r2_p20=r2s, nsteps=i + 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        return s / C  # :73 (transposed)


# --------------------------------------------------------------------------------------------
# QC (batch kurtosis)
# --------------------------------------------------------------------------------------------
def batch_kurtosis(X, batches):
    """_nam.py:78-82.  X is N x n (cells x samples), ``batches`` has one label per column of X.
    Per cell: Pearson kurtosis (Fisher + 3) across the per-batch means."""
    batches = np.asarray(batches)
    means = np.stack([X[:, batches == b].mean(axis=1) for b in np.unique(batches)], axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        return st.kurtosis(means, axis=1) + 3


def qc_keep(X, batches):
    """_nam.py:85-99.  Keep mask over cells; all True when there is a single batch."""
    if len(np.unique(batches)) == 1:  # :89
        return np.ones(X.shape[0], dtype=bool), None
    kurt = batch_kurtosis(X, batches)
    threshold = max(6, 2 * np.median(kurt))  # :94  (python max: NaN median -> 6)
    return kurt < threshold, threshold  # :96


def nam(data, sid_name, batches=None, nsteps=None, self_weight=1, diagnostics=None,
        faithful=False, **kwargs):
    """_nam.py:179-193.  Returns (DataFrame samples x kept cells, keep mask) like the reference."""
    sid = data.obs[sid_name]
    labels, codes = sample_codes(sid)
    if batches is None:  # :185-186
        batches = pd.Series(np.ones(len(sid.unique())), index=sid.unique())
    A = get_connectivity(data)
    X = nam_raw(A, codes, len(labels), nsteps=nsteps, self_weight=self_weight,
                diagnostics=diagnostics, faithful=faithful)
    b = batches.reindex(labels).to_numpy()  # _nam.py:79
    keep, thr = qc_keep(X, b)
    if diagnostics is not None:
        diagnostics["qc_threshold"] = thr
    out = pd.DataFrame(X[keep].T, index=labels, columns=data.obs.index[keep], dtype=float)
    out.index.name = sid_name  # :74
    return out, keep


# --------------------------------------------------------------------------------------------
# residualisation + SVD (kernels ii, iii)
# --------------------------------------------------------------------------------------------
def _std1(a, axis=0):
    return a.std(axis=axis, ddof=1)


def design_matrix(covs, batches):
    """_nam.py:123-139.  Returns (C [n x r], number of batch columns).  pandas ``.std`` is ddof=1."""
    n = len(batches) if batches is not None else len(covs)
    if covs is None:
        cov = np.ones((n, 0))
    else:
        cov = np.asarray(covs, dtype=np.float64).reshape(n, -1)
        cov = (cov - cov.mean(axis=0)) / _std1(cov)  # :126
    if batches is None or len(np.unique(batches)) == 1:  # :128
        return cov, 0
    ub = np.unique(batches)  # get_dummies column order == sorted unique, :137
    B = (np.asarray(batches)[:, None] == ub[None, :]).astype(np.float64)
    B = (B - B.mean(axis=0)) / _std1(B)  # :138
    return np.concatenate([B, cov], axis=1), B.shape[1]


def projector_stage(C, nb, ridge):
    """_nam.py:145-146 (or :133 when nb == 0).  Returns W [r x n] with M = I - C.W."""
    n, r = C.shape
    CtC = C.T.dot(C)
    if nb > 0:
        L = np.diag([1.0] * nb + [0.0] * (r - nb))
        CtC = CtC + ridge * n * L
    return np.linalg.solve(CtC, C.T)


def svd_gram(X):
    """_nam.py:102-106 on an already centred/standardised N x n matrix.
    Returns (U [n x n], svs [n] = squared singular values, G)."""
    G = X.T.dot(X)
    U, svs, _ = np.linalg.svd(G)  # :105
    return U, svs, G


def svd_nam(NAM):
    """_nam.py:102-115 — public; takes the reference's samples x cells DataFrame."""
    X = NAM.to_numpy(dtype=np.float64).T
    X = X - X.mean(axis=1, keepdims=True)  # :103
    X = X / _std1(X, axis=1)[:, None]  # :104
    U, svs, _ = svd_gram(X)
    with np.errstate(divide="ignore", invalid="ignore"):
        V = X.dot(U) / np.sqrt(svs)  # :106
    pcs = ["PC" + str(i) for i in range(1, len(svs) + 1)]
    return (pd.DataFrame(U, index=NAM.index, columns=pcs), pd.Series(svs, index=pcs),
            pd.DataFrame(V, index=NAM.columns, columns=pcs))


def resid_nam(X, covs, batches, ridges=None, npcs=None, diagnostics=None, faithful=False):
    """_nam.py:118-177 in N x n layout.  Returns Namespace(M, r, X (= namresid^T), U, svs, ...)."""
    N, n = X.shape
    X = X - X.mean(axis=1, keepdims=True)  # :122
    C, nb = design_matrix(covs, batches)
    r = C.shape[1]
    M = np.eye(n)
    ridge_log = []
    if nb == 0:
        if r > 0:  # :133
            M = np.eye(n) - C.dot(projector_stage(C, 0, 0.0))
            X = X.dot(M.T) if faithful else X - X.dot(projector_stage(C, 0, 0.0).T).dot(C.T)
    else:
        for ridge in (DEFAULT_RIDGES if ridges is None else ridges):  # :141-144
            W = projector_stage(C, nb, ridge)
            M = np.eye(n) - C.dot(W)  # :146  (only the last M survives, :169)
            X = X.dot(M.T) if faithful else X - X.dot(W.T).dot(C.T)  # :148, cumulative
            med = np.median(batch_kurtosis(X, batches))  # :150-155
            ridge_log.append((ridge, med))
            if med <= 6:
                break
    X = X / _std1(X, axis=1)[:, None]  # :159
    U, svs, G = svd_gram(X)  # :163 (re-centre / re-standardise at :103-104 are no-ops here)
    res = Namespace(M=M, r=r, X=X, U=U, svs_full=svs, G=G)
    if npcs is None:
        npcs = n
    res.svs = svs[:npcs]  # :174
    res.varexp = svs / n / N  # :175
    if diagnostics is not None:
        diagnostics["ridge_log"] = ridge_log
    return res


# --------------------------------------------------------------------------------------------
# permutations + tail counts (_stats.py)
# --------------------------------------------------------------------------------------------
def conditional_permutation_indices(B, num):
    """_stats.py:4-16 — index matrix only (n x num).  Consumes the *global* numpy legacy RNG in the
    reference's order: one ``randn(len(batch), num)`` block per batch, batches in sorted order."""
    B = np.asarray(B)
    batchind = [np.where(B == b)[0] for b in np.unique(B)]
    ix = np.concatenate([bi[np.argsort(np.random.randn(len(bi), num), axis=0)] for bi in batchind])
    bix = np.zeros((len(B), num), dtype=np.int64)
    bix[np.concatenate(batchind)] = ix
    return bix


def conditional_permutation(B, Y, num):
    """_stats.py:4-18."""
    return np.asarray(Y)[conditional_permutation_indices(B, num)]


def grouplevel_permutation(G, Y, num):
    """_stats.py:20-32."""
    G = np.asarray(G)
    Y = np.asarray(Y)
    Gu = np.unique(G)
    Yg = np.array([Y[G == g][0] for g in Gu])
    Gind = np.searchsorted(Gu, G)
    if (Yg[Gind] != Y).any():
        print("ERROR: the value of Y is not identical within each group of samples")
        return None
    ix = np.argsort(np.random.randn(len(Yg), num), axis=0)
    return Yg[ix][Gind]


def threshold_edges(t, atol=1e-8, rtol=1e-5):
    """_stats.py:51 — histogram edges for (ascending) thresholds t: t^2 - atol - rtol.t^2."""
    t2 = np.asarray(t, dtype=np.float64) ** 2
    return t2 - atol - rtol * t2


def tail_counts(t, znull, faithful=False):
    """_stats.py:34-62 restricted to ascending ``t`` (the only way the path calls it).
    tails[k, i] = #{cells : znull[cell, k]^2 >= edge_i}."""
    if znull.ndim == 1:
        znull = znull[:, None]
    edges = threshold_edges(t)
    if faithful:
        bins = np.concatenate([edges, [np.inf]])
        hist = np.array([np.histogram(z2, bins=bins)[0] for z2 in znull.T ** 2])
        tails = np.flip(hist, axis=1)
        np.cumsum(tails, axis=1, out=tails)
        return np.flip(tails, axis=1)
    out = np.empty((znull.shape[1], len(edges)), dtype=np.int64)
    for k in range(znull.shape[1]):
        z2 = np.sort(znull[:, k] ** 2)
        out[k] = len(z2) - np.searchsorted(z2, edges, side="left")
    return out


def empirical_fdrs(z, znull, thresholds, faithful=False):
    """_stats.py:64-83."""
    tails = tail_counts(thresholds, znull, faithful=faithful)
    ranks = tail_counts(thresholds, z, faithful=faithful)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (tails / ranks).mean(axis=0)


# --------------------------------------------------------------------------------------------
# association (_association.py)
# --------------------------------------------------------------------------------------------
def default_ks(n):
    """_association.py:25-28."""
    incr = max(int(0.02 * n), 1)
    maxnpcs = max(min(4 * incr, int(n / 5)), 1)
    return np.arange(incr, maxnpcs + 1, incr)


def minp_stats_matrix(Z, M, U, ks, n, r):
    """_association.py:35-61 for every column of Z (n x K) at once.
    Returns (k, p, r2) arrays of length K.  ``zcond.std()`` at :52 is a pandas Series -> ddof=1."""
    Zc = M.dot(Z)
    Zc = Zc / _std1(Zc, axis=0)
    ssered = (Zc * Zc).sum(axis=0)
    kmax = int(max(ks))
    B = U[:, :kmax].T.dot(Zc)
    ps = np.empty((len(ks), Z.shape[1]))
    r2s = np.empty_like(ps)
    for a, k in enumerate(ks):
        # the reference forms yhat explicitly and sums squared residuals (:42); U is orthonormal
        # so this equals ssered - |beta|^2 up to rounding.  Keep the explicit form for fidelity.
        Zhat = U[:, :k].dot(B[:k])
        ssefull = ((Zhat - Zc) ** 2).sum(axis=0)
        with np.errstate(divide="ignore", invalid="ignore"):
            f = ((ssered - ssefull) / k) / (ssefull / n)  # :45 (divides by n, not dof)
            ps[a] = st.f.sf(f, k, n - (1 + r + k))  # :46
            r2s[a] = 1 - ssefull / ssered  # :47
    with np.errstate(invalid="ignore"):
        pick = np.nanargmin(ps, axis=0)  # :60
    cols = np.arange(Z.shape[1])
    return np.asarray(ks)[pick], ps[pick, cols], r2s[pick, cols]


def _minp_stats_single(z, M, U, ks, n, r):
    """_association.py:50-61, one phenotype vector — the reference's per-permutation call."""
    zc = M.dot(z)
    zc = zc / zc.std(ddof=1)
    ps, r2s = [], []
    for k in ks:
        Xpc = U[:, :k]
        zhat = Xpc.dot(Xpc.T.dot(zc))
        ssefull = (zhat - zc).dot(zhat - zc)
        ssered = zc.dot(zc)
        f = ((ssered - ssefull) / k) / (ssefull / n)
        ps.append(st.f.sf(f, k, n - (1 + r + k)))
        r2s.append(1 - ssefull / ssered)
    k_ = np.nanargmin(ps)
    return ks[k_], ps[k_], r2s[k_]


def association_core(U, X, M, r, y, batches, donorids, ks=None, Nnull=1000,
                     force_permute_all=False, local_test=True, seed=None, faithful=False):
    """_association.py:10-129.  X is N' x n (namresid transposed); y, batches, donorids are
    length-n arrays already filtered and ordered like the columns of X."""
    if seed is not None:
        np.random.seed(seed)  # :15-16
    if force_permute_all:
        batches = np.ones(len(y))  # :17-18
    y = np.asarray(y, dtype=np.float64)
    y = (y - y.mean()) / y.std()  # :22  (ndarray -> ddof=0)
    n = len(y)
    if ks is None:
        ks = default_ks(n)
    if max(ks) + r >= n:  # :29-33
        raise ValueError(
            "Maximum number of PCs plus number of covariates must be less than n-1. "
            f"Currently it is {max(ks) + r} while n is {n}. Either reduce the number of covariates "
            "or reduce the number of PCs to consider using the optional argument ks=[...].")

    k, p, r2 = (a[0] for a in minp_stats_matrix(y[:, None], M, U, ks, n, r))  # :64
    if k == max(ks):  # :65-67
        warnings.warn(("data supported use of {} NAM PCs, which is the maximum considered. "
                       'Consider allowing more PCs by using the "ks" argument.').format(k))
    ycond = M.dot(y)
    ycond = ycond / ycond.std(ddof=1)  # :70-71 (Series.std -> ddof=1 in the reference)
    beta = U[:, :k].T.dot(ycond)  # :72
    yhat = U[:, :k].dot(beta)
    r2_perpc = (beta / np.sqrt(ycond.dot(ycond))) ** 2  # :74

    # :77 — raw standardised y, not ycond.  The reference averages a samples x cells DataFrame
    # down the sample axis, i.e. numpy adds the n rows one after another; the summation order is
    # kept because len(np.arange(m/4, m, m/400)) at :102 is 300 or 301 depending on the last bit
    # of m = max|ncorrs|.
    ncorrs = (y[:, None] * np.ascontiguousarray(X.T)).sum(axis=0) / n

    if donorids is not None:  # :80-83
        y_ = grouplevel_permutation(donorids, y, Nnull)
    else:
        y_ = conditional_permutation(batches, y, Nnull)
    if faithful:  # :84 — python loop over permutations
        stats = np.array([_minp_stats_single(col, M, U, ks, n, r)[1:] for col in y_.T])
        nullminps, nullr2s = stats.T
    else:
        _, nullminps, nullr2s = minp_stats_matrix(y_, M, U, ks, n, r)
    nhit = int((nullminps <= p + 1e-8).sum())
    pfinal = (nhit + 1) / (Nnull + 1)  # :85
    if nhit == 0:  # :86-88
        warnings.warn("global association p-value attained minimal possible value. "
                      "Consider increasing Nnull")

    fdrs, fdr_5p_t, fdr_10p_t = None, None, None
    if local_test:  # :92-120
        Kl = min(1000, Nnull)
        ycond_ = M.dot(y_[:, :Kl])
        ycond_ = ycond_ / _std1(ycond_, axis=0)  # :97 — M is a DataFrame, so .std is ddof=1
        maxcorr = max(np.abs(ncorrs).max(), 0.001)  # :101
        thresholds = np.arange(maxcorr / 4, maxcorr, maxcorr / 400)  # :102
        if faithful:
            nullncorrs = np.abs(X.dot(ycond_) / n)  # :99, N' x Kl materialised like the reference
            fdr_vals = empirical_fdrs(ncorrs, nullncorrs, thresholds, faithful=True)
            del nullncorrs
        else:
            edges = threshold_edges(thresholds)
            tails = np.empty((Kl, len(edges)), dtype=np.int64)
            for c0 in range(0, Kl, 64):  # blocked so N' x Kl is never materialised
                z2 = np.sort((X.dot(ycond_[:, c0:c0 + 64]) / n) ** 2, axis=0)
                for j in range(z2.shape[1]):
                    tails[c0 + j] = z2.shape[0] - np.searchsorted(z2[:, j], edges, side="left")
            ranks = tail_counts(thresholds, ncorrs)
            with np.errstate(divide="ignore", invalid="ignore"):
                fdr_vals = (tails / ranks).mean(axis=0)
        absn = np.abs(ncorrs)
        if faithful:
            num_detected = [(absn > t).sum() for t in thresholds]  # :108
        else:
            srt = np.sort(absn)
            num_detected = len(srt) - np.searchsorted(srt, thresholds, side="right")
        fdrs = pd.DataFrame({"threshold": thresholds, "fdr": fdr_vals,
                             "num_detected": np.asarray(num_detected)})
        if not np.min(fdrs.fdr) > 0.05:  # :111-114
            fdr_5p_t = fdrs[fdrs.fdr <= 0.05].iloc[0].threshold
        if not np.min(fdrs.fdr) > 0.1:  # :115-118
            fdr_10p_t = fdrs[fdrs.fdr <= 0.1].iloc[0].threshold

    return Namespace(p=pfinal, nullminps=nullminps, k=k, ncorrs=ncorrs, fdrs=fdrs,
                     fdr_5p_t=fdr_5p_t, fdr_10p_t=fdr_10p_t, yresid_hat=yhat, yresid=ycond,
                     ks=ks, beta=beta, r2=r2, r2_perpc=r2_perpc,
                     nullr2_mean=nullr2s.mean(), nullr2_std=nullr2s.std())


def check_inputs(data, y, sid_name, batches, covs, donorids, allow_low_sample_size):
    """_association.py:131-173 — same exception types and messages."""
    if not isinstance(y, pd.Series):
        raise TypeError(f"'y' must be a pandas Series, but got {type(y)}")
    if batches is not None and not isinstance(batches, pd.Series):
        raise TypeError(f"'batches' must be a pandas Series, but got {type(batches)}")
    if covs is not None and not isinstance(covs, pd.DataFrame):
        raise TypeError(f"'covs' must be a pandas DataFrame, but got {type(covs)}")
    if donorids is not None and not isinstance(donorids, pd.Series):
        raise TypeError(f"'donorids' must be a pandas Series, but got {type(donorids)}")
    present = data.obs[sid_name].unique()
    if not set(y.index).issubset(set(present)):
        print("WARNING: index of 'y' contains values not present in 'data[sid_name]'. "
              "These samples will be ignored.")
    if not set(present).issubset(set(y.index)):
        raise ValueError("'data[sid_name]' contains values not present in the index of 'y'.")
    if batches is not None and donorids is not None:
        raise ValueError("We do not currently support conditioning on batch "
                         "while also accounting for multiple samples per donor")
    if batches is None:
        batches = pd.Series(np.ones(len(y)), index=y.index)
    if covs is not None:
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


def cell_fdr_lookup(coef, fdrs, faithful=False):
    """_association.py:234-237 — per cell, the smallest fdr among thresholds <= |coef| (else 1)."""
    if faithful:
        def min_fdr_for_corr(c):
            m = fdrs.loc[fdrs.threshold <= abs(c)].fdr
            return m.min() if not m.empty else 1
        return pd.Series(coef).apply(min_fdr_for_corr).to_numpy()
    t = fdrs.threshold.to_numpy()
    pm = np.fmin.accumulate(fdrs.fdr.to_numpy())
    idx = np.searchsorted(t, np.nan_to_num(np.abs(coef), nan=-1.0), side="right")
    return np.where(idx > 0, pm[np.maximum(idx - 1, 0)], 1.0)


def association(data, y, sid_name, batches=None, covs=None, donorids=None, ks=None,
                key_added="coef", max_frac_pcs=0.15, nsteps=None, show_progress=False,
                allow_low_sample_size=False, return_full=False, ridges=None, faithful=False,
                diagnostics=None, **kwargs):
    """_association.py:193-242.  Same signature, side effects on ``data.obs`` and return value."""
    batches, filter_samples = check_inputs(data, y, sid_name, batches, covs, donorids,
                                           allow_low_sample_size)
    # ---- compute_nam_and_reindex (:175-191) ----
    NAM, kept = nam(data, sid_name, batches=batches, nsteps=nsteps, diagnostics=diagnostics,
                    faithful=faithful)
    NAM = NAM.reindex(y.index)  # :178
    fs = filter_samples.to_numpy()
    X = NAM.to_numpy()[fs].T  # N_qc x n  (:181)
    zero_var = np.where(_std1(X, axis=1) == 0)[0]  # :182
    nz = np.flatnonzero(kept)
    kept[nz[zero_var]] = False  # :183-184
    X = np.delete(X, zero_var, axis=0)  # :185
    batches = batches.reindex(y.index)
    covs = covs.reindex(y.index) if covs is not None else None
    donorids = donorids.reindex(y.index) if donorids is not None else None

    n = int(fs.sum())
    npcs = min(n, max([10] + [int(max_frac_pcs * n)] + [ks if ks is not None else []][0]))  # :207
    res = resid_nam(X, covs[fs].to_numpy() if covs is not None else None,
                    batches[fs].to_numpy(), ridges=ridges, npcs=npcs, diagnostics=diagnostics,
                    faithful=faithful)
    core = association_core(res.U, res.X, res.M, res.r, y[fs].to_numpy(), batches[fs].to_numpy(),
                            donorids[fs].to_numpy() if donorids is not None else None,
                            ks=ks, faithful=faithful, **kwargs)

    # ---- assemble the reference's Namespace (:223-225; _nam.py:168-175) ----
    sids = y.index[fs]
    cells = data.obs.index[kept]
    pcs = ["PC" + str(i) for i in range(1, n + 1)]
    out = Namespace()
    out.M = pd.DataFrame(res.M, index=sids, columns=sids)
    out.r = res.r
    out.namresid = pd.DataFrame(res.X.T, index=sids, columns=cells)
    out.namresid_sampleXpc = pd.DataFrame(res.U, index=sids, columns=pcs)
    if return_full:
        with np.errstate(divide="ignore", invalid="ignore"):
            V = res.X.dot(res.U) / np.sqrt(res.svs_full)  # _nam.py:106
        out.namresid_nbhdXpc = pd.DataFrame(V, index=cells, columns=pcs)
    out.namresid_svs = pd.Series(res.svs_full, index=pcs)[:npcs]
    out.namresid_varexp = pd.Series(res.varexp, index=pcs)
    out.__dict__.update(vars(core))
    out.ncorrs = pd.Series(core.ncorrs, index=cells)
    out.yresid = pd.Series(core.yresid, index=sids)
    out.nam = pd.DataFrame(X.T, index=sids, columns=cells)
    out.kept = kept
    out.G = res.G

    # ---- obs write-back (:228-237) ----
    if key_added in data.obs:
        warnings.warn(f"Key '{key_added}' already exists in data.obs. Overwriting.")
    coef = np.full(len(data.obs), np.nan)
    coef[kept] = core.ncorrs
    data.obs[key_added] = coef
    if core.fdrs is not None:
        data.obs[f"{key_added}_fdr"] = cell_fdr_lookup(coef, core.fdrs, faithful=faithful)
    return out if return_full else out.p
