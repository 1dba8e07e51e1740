"""Shared comparison helpers for the parity tests."""
import json
import os
import warnings

import numpy as np

from tests.golden import cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(stem):
    arrays = dict(np.load(os.path.join(GOLD, stem + ".npz")))
    scalars = json.load(open(os.path.join(GOLD, stem + ".json")))
    return arrays, scalars


def synth_raw(arrays, name):
    pre = name + "/raw_"
    return {k[len(pre):]: v for k, v in arrays.items() if k.startswith(pre)}


def run_association(fn, data, kwargs, np_seed=None, **extra):
    """Call an ``association`` implementation the way the golden cases do; returns (res, warnings)."""
    if np_seed is not None:
        np.random.seed(np_seed)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        res = fn(data, return_full=True, **kwargs, **extra)
    warns = sorted({str(x.message)[:60] for x in w if issubclass(x.category, UserWarning)})
    return res, warns


def sign_align(a, b):
    """Flip the columns of ``a`` to the sign of the matching columns of ``b``."""
    s = np.sign((a * b).sum(axis=0))
    s[s == 0] = 1
    return a * s


def _knife_edge_counts(ncorrs, cuts, tol):
    """Number of reference cells whose |coefficient| lies within ``tol`` of each cut."""
    a = np.sort(np.abs(ncorrs))
    return np.searchsorted(a, cuts + tol, side="right") - np.searchsorted(a, cuts - tol, side="left")


def assert_matches_golden(res, data, key, arrays, scalars, name, warns=None, rtol=1e-9, atol=1e-12,
                          exact_sets=True, check_full=True, fp32=False):
    """Compare a result Namespace + obs columns with the reference outputs stored for ``name``.

    ``rtol`` is the relative tolerance for floating-point outputs (1e-9 for the float64 oracle,
    1e-5 — the north-star tolerance — for the fp32 CUDA path); integer / index outputs are exact.

    ``fp32=True`` (the CUDA path, whose diffusion state is fp32): threshold counts may differ from
    the reference only by cells whose coefficient lies within fp32 rounding distance (1e-6) of that
    threshold — the same knife edge that makes the reference's own counts depend on its BLAS
    summation order, three decimal digits wider — and quantities that depend on individual
    eigenvectors (beta, yhat, U, V: conditioned by the eigen-gap, not by the input error) get a
    1e-3 tolerance.  Everything the north star lists (kept set, p, k, singular values, coefficients,
    FDR-passing set) keeps the strict bar.
    """
    vec_rtol = 1e-3 if fp32 else rtol
    ref_nc, ref_fd = arrays[name + "/ncorrs"], arrays[name + "/fdrs"]
    # coefficients are only required to rtol * max|coef| (checked below), so a threshold count may
    # move by the number of reference cells that close to the threshold
    edge_tol = rtol * np.abs(ref_nc).max() if fp32 else 0.0
    sc = scalars[name]
    g = lambda f: arrays[name + "/" + f]  # noqa: E731
    # ---- integer / index outputs: exact ----
    assert list(map(int, res.ks)) == sc["ks"]
    assert int(res.k) == int(sc["k"])
    assert int(res.r) == int(sc["r"])
    np.testing.assert_array_equal(np.asarray(res.kept), g("kept"))
    assert float(res.p) == sc["p"], (res.p, sc["p"])  # a count ratio
    if warns is not None:
        assert warns == sc["warnings"], warns
    coef_fdr = data.obs[key + "_fdr"].to_numpy()
    if exact_sets:
        assert int((coef_fdr <= 0.05).sum()) == sc["n_fdr05"]
        assert int((coef_fdr <= 0.10).sum()) == sc["n_fdr10"]
        np.testing.assert_array_equal(coef_fdr <= 0.05, g("coef_fdr") <= 0.05)
        nt = min(len(g("fdrs")), len(res.fdrs))
        slack = _knife_edge_counts(ref_nc, ref_fd[:nt, 0], edge_tol) if fp32 else np.zeros(nt)
        nd_err = np.abs(g("fdrs")[:nt, 2] - res.fdrs["num_detected"].to_numpy()[:nt])
        assert (nd_err <= slack).all(), (nd_err.max(), np.flatnonzero(nd_err > slack))
    # ---- floating-point outputs ----
    close = lambda a, b, **kw: np.testing.assert_allclose(  # noqa: E731
        np.asarray(a, dtype=np.float64), b, rtol=kw.get("rtol", rtol), atol=kw.get("atol", atol))
    close(res.ncorrs.to_numpy(), g("ncorrs"), atol=max(atol, rtol * np.abs(g("ncorrs")).max()))
    close(data.obs[key].to_numpy(), g("coef"), atol=max(atol, rtol * np.nanmax(np.abs(g("coef")))))
    svs = g("svs")
    close(res.namresid_svs.to_numpy(), svs, atol=max(atol, rtol * 1e-3 * svs.max()))
    close(res.r2, sc["r2"])
    kk = int(sc["k"])
    close(np.abs(res.beta), np.abs(g("beta")), rtol=vec_rtol, atol=max(atol, vec_rtol * np.abs(g("beta")).max()))
    close(res.r2_perpc, g("r2_perpc"), rtol=vec_rtol, atol=max(atol, vec_rtol))
    close(res.yresid.to_numpy(), g("yresid"), atol=max(atol, rtol))
    close(res.yresid_hat, g("yresid_hat"), rtol=vec_rtol, atol=max(atol, vec_rtol))
    close(res.nullminps, g("nullminps"), rtol=max(rtol, 1e-9) * 50, atol=1e-300)
    close(res.nullr2_mean, sc["nullr2_mean"], rtol=max(rtol, 1e-9) * 10)
    close(res.nullr2_std, sc["nullr2_std"], rtol=max(rtol, 1e-9) * 10)
    # len(np.arange(m/4, m, m/400)) (_association.py:102) is 300 or 301 depending on the last
    # bit of m = max|ncorrs| — a floating-point knife-edge of the reference itself (its own result
    # changes with the BLAS summation order).  The common 300 rows must agree.
    fd = g("fdrs")
    assert len(fd) in (300, 301) and len(res.fdrs) in (300, 301)
    nt = min(len(fd), len(res.fdrs))
    close(res.fdrs["threshold"].to_numpy()[:nt], fd[:nt, 0])
    fdr_tol = np.full(nt, max(atol, rtol))
    if fp32:
        # fdr_i = mean_k(tails_ki) / ranks_i (_stats.py:79-80).  One knife-edge cell moving across
        # edge i changes ranks_i by one (relative 1/ranks_i); a null value doing the same changes
        # mean_k(tails_ki) by 1/Kl (the null matrix is not stored in the fixtures, so allow three).
        cuts = np.sqrt(np.maximum(fd[:nt, 0] ** 2 * (1 - 1e-5) - 1e-8, 0))  # _stats.py:51
        ranks = np.maximum((np.abs(ref_nc)[:, None] >= cuts[None, :]).sum(0), 1)
        n_local = min(1000, len(g("nullminps")))
        rel = _knife_edge_counts(ref_nc, cuts, edge_tol) / ranks + 1e-4
        fdr_tol = rel * np.abs(fd[:nt, 1]) + 3.0 / (n_local * ranks) + max(atol, rtol)
    err = np.abs(res.fdrs["fdr"].to_numpy()[:nt] - fd[:nt, 1])
    assert (err <= fdr_tol).all(), (err.max(), np.flatnonzero(err > fdr_tol))
    for f in ("fdr_5p_t", "fdr_10p_t"):
        if sc[f] is None:
            assert getattr(res, f) is None
        else:
            close(getattr(res, f), sc[f])
    if fp32:
        # per-cell fdr = running minimum of fdr over thresholds <= |coef| (_association.py:234-237):
        # a cell within edge_tol of a threshold may land on either side of it
        t = fd[:nt, 0]
        pm = np.concatenate([[1.0], np.fmin.accumulate(fd[:nt, 1])])
        tol_cum = np.concatenate([[0.0], np.maximum.accumulate(fdr_tol)])
        c = np.nan_to_num(np.abs(g("coef")), nan=-1.0)
        lo = np.searchsorted(t, c - edge_tol, side="right")
        hi = np.searchsorted(t, c + edge_tol, side="right")
        assert ((coef_fdr <= pm[lo] + tol_cum[hi]) & (coef_fdr >= pm[hi] - tol_cum[hi])).all()
    else:
        close(coef_fdr, g("coef_fdr"), atol=max(atol, rtol))
    close(res.M.to_numpy(), g("M"), atol=max(atol, rtol))
    if check_full:
        n_top = max(kk, 4)
        U = sign_align(res.namresid_sampleXpc.to_numpy()[:, :n_top], g("U")[:, :n_top])
        close(U, g("U")[:, :n_top], atol=max(1e-9, vec_rtol * 20))
        close(res.namresid_varexp.to_numpy()[:n_top], g("varexp")[:n_top])
        close(res.namresid.to_numpy()[:, :64], g("namresid_head"), atol=max(atol, rtol * 10))
        nh = g("nam_head")
        close(res.nam.to_numpy()[:, :64], nh, atol=max(atol, rtol * np.abs(nh).max()))
        V = sign_align(res.namresid_nbhdXpc.to_numpy()[:64, :n_top], g("V_head")[:, :n_top])
        close(V, g("V_head")[:, :n_top], atol=max(1e-9, vec_rtol * 20 * np.abs(g("V_head")).max()))
