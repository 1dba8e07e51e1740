"""GPU parity: the CUDA path (through the C-ABI / public ``cna_b200.tl`` surface) against
(a) the committed outputs of the unmodified reference (tests/golden) and (b) the CPU oracle on
seeded inputs.  Integer / index outputs must be exact; floating-point outputs within the
north-star tolerance of 1e-5 relative (the device state is fp32, the reference fp64).
"""
import os
import warnings

import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp

from tests import helpers
from tests.golden import cases

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # BASELINE.json north_star: singular values, coefficients, p within 1e-5 relative


@pytest.fixture(scope="module")
def cna():
    import cna_b200
    from cna_b200 import _lib
    _lib.load()
    return cna_b200


@pytest.fixture(scope="module")
def demo():
    return helpers.load_golden("demo_cases")


@pytest.fixture(scope="module")
def synth():
    return helpers.load_golden("synth_cases")


def test_library_loaded_and_counts_launches(cna):
    import torch
    from cna_b200 import _lib
    before = _lib.launch_count()
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    _lib.absmax(torch.tensor([1.0, -3.0, 2.0], dtype=torch.float64, device="cuda"), None, out)
    assert out.item() == 3.0
    assert _lib.launch_count() == before + 1


@pytest.mark.parametrize("name", list(cases.DEMO_CASES))
def test_demo_cases_match_reference(cna, demo, name):
    arrays, scalars = demo
    spec = cases.DEMO_CASES[name]
    data, kwargs = cases.build_demo_case(cases.load_demo_graph(), spec)
    res, warns = helpers.run_association(cna.tl.association, data, kwargs, spec.get("np_seed"))
    helpers.assert_matches_golden(res, data, kwargs.get("key_added", "coef"), arrays, scalars, name,
                                  warns=warns, rtol=RTOL, atol=1e-9, fp32=True)


@pytest.mark.parametrize("name", list(cases.SYNTH_CASES))
def test_synth_cases_match_reference(cna, synth, name):
    arrays, scalars = synth
    spec = cases.SYNTH_CASES[name]
    data, kwargs = cases.build_synth_case(spec, helpers.synth_raw(arrays, name))
    res, warns = helpers.run_association(cna.tl.association, data, kwargs)
    helpers.assert_matches_golden(res, data, "coef", arrays, scalars, name, warns=warns, rtol=RTOL,
                                  atol=1e-9, fp32=True)


def test_notebook_numbers(cna):
    """demo/demo.ipynb cell 10: p = 0.000999000999000999 and 9555 neighbourhoods at FDR 5 %."""
    data, kwargs = cases.build_demo_case(cases.load_demo_graph(), cases.DEMO_CASES["case_male_batch"])
    np.random.seed(0)
    with pytest.warns(UserWarning, match="minimal possible value"):
        p = cna.tl.association(data, **kwargs)
    assert p == 0.000999000999000999
    assert int((data.obs["case_coef_fdr"] <= 0.05).sum()) == 9555
    # resident graph handle: same numbers, no re-upload
    data2, kwargs2 = cases.build_demo_case(cases.load_demo_graph(), cases.DEMO_CASES["male_case_batch"])
    h = cna.tl.to_device(data2)
    np.random.seed(0)
    with pytest.warns(UserWarning):
        p2 = cna.tl.association(h, **kwargs2)
    assert p2 == 0.000999000999000999
    assert int((data2.obs["male_coef_fdr"] <= 0.05).sum()) == 4509


def test_nam_svd_diffuse_match_reference(cna, demo):
    arrays, _ = demo
    data = cases.demo_anndata()
    meta = cases.demo_sample_meta()
    NAM, keep = cna.tl.nam(data, "id", batches=meta.batch)
    assert NAM.shape == (50, 10000) and NAM.index.name == "id"
    np.testing.assert_array_equal(keep, arrays["nam/keep"])
    np.testing.assert_allclose(NAM.to_numpy()[:, :256], arrays["nam/NAM_head"], rtol=RTOL, atol=1e-10)
    np.testing.assert_allclose(NAM.to_numpy().sum(axis=1), arrays["nam/rowsum"], rtol=RTOL)
    U, svs, V = cna.tl.svd_nam(NAM)
    top = arrays["nam/svs"][:-2]  # the last 1-2 singular values are ~0 (null space)
    np.testing.assert_allclose(svs.to_numpy()[:len(top)], top, rtol=RTOL, atol=RTOL * 1e-3 * top.max())
    np.testing.assert_allclose(helpers.sign_align(U.to_numpy()[:, :5], arrays["nam/U"][:, :5]),
                               arrays["nam/U"][:, :5], atol=2e-5)
    np.testing.assert_allclose(helpers.sign_align(V.to_numpy()[:64, :4], arrays["nam/V_head"][:, :4]),
                               arrays["nam/V_head"][:, :4], atol=2e-5 * np.abs(arrays["nam/V_head"]).max())
    for s in (1, 2, 3):
        got = cna.tl.nam(data, "id", nsteps=s)[0].to_numpy()[:, :256]
        np.testing.assert_allclose(got, arrays[f"nam/steps{s}_head"], rtol=RTOL, atol=1e-10)
    # public diffuse() on user vectors runs in float64 on the device
    np.testing.assert_allclose(cna.tl.diffuse(data, arrays["diffuse/s0"], 2), arrays["diffuse/s2"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(cna.tl.diffuse(data, arrays["diffuse/s0"], 3, self_weight=0.5),
                               arrays["diffuse/s3_w05"], rtol=1e-12, atol=1e-15)
    steps = list(cna.tl.diffuse_stepwise(data, arrays["diffuse/s0"], maxnsteps=2))
    assert len(steps) == 2
    np.testing.assert_allclose(steps[1], arrays["diffuse/s2"], rtol=1e-12, atol=1e-15)


def test_auto_stop_diagnostics(cna):
    """the demo auto-stops after 4 steps with these median kurtoses."""
    from cna_b200.tl import _nam
    data = cases.demo_anndata()
    st = _nam._nam_device(data, "id")
    assert st.nsteps == 4
    np.testing.assert_allclose(st.medkurt, [16.40326205754668, 12.757992457504468,
                                            6.115484918150449, 3.177490071480447], rtol=RTOL)


def test_input_errors(cna):
    data = cases.demo_anndata()
    meta = cases.demo_sample_meta()
    with pytest.raises(TypeError):
        cna.tl.association(data, meta.case.to_numpy(), "id")
    with pytest.raises(TypeError):
        cna.tl.association(data, meta.case, "id", covs=meta.male)
    with pytest.raises(ValueError):
        cna.tl.association(data, meta.case.iloc[:40], "id")
    with pytest.raises(ValueError):
        cna.tl.association(data, meta.case, "id", batches=meta.batch, donorids=meta.batch)
    y = meta.case.copy()
    y.iloc[:45] = np.nan
    with pytest.raises(ValueError, match="fewer than 10 samples"):
        cna.tl.association(data, y, "id")
    with pytest.raises(ValueError, match="Maximum number of PCs"):
        cna.tl.association(data, meta.case, "id", ks=[50], nsteps=1, Nnull=10)
    with pytest.raises(TypeError):
        cna.tl.association(data, meta.case, "id", self_weight=2)


# ---------------------------------------------------------------------------------------------
# kernel-level parity against the oracle on seeded inputs (C-ABI wrappers called directly)
# ---------------------------------------------------------------------------------------------
def _random_graph(n, deg, seed, hub=None):
    rng = np.random.default_rng(seed)
    rows = np.repeat(np.arange(n), deg)
    cols = rng.integers(0, n, size=n * deg)
    if hub is not None:  # one heavy row / column, and an isolated cell (empty row)
        cols[: n // 3] = hub
    A = sp.csr_matrix((rng.uniform(0.05, 1.0, n * deg), (rows, cols)), shape=(n, n))
    A = A.maximum(A.T).tolil()
    A.setdiag(0)
    A[n - 1, :] = 0
    A[:, n - 1] = 0
    A = A.tocsr()
    A.eliminate_zeros()
    A.sort_indices()
    return A


@pytest.mark.parametrize("n,S,deg", [(3000, 50, 7), (2048, 200, 12), (1500, 100, 5), (700, 333, 9), (257, 3, 4)])
def test_diffusion_kernels_vs_oracle(cna, n, S, deg):
    import torch
    from cna_b200 import _lib
    from cna_b200.tl._graph import DeviceGraph
    from oracle import cna_oracle as orc
    A = _random_graph(n, deg, seed=n + S, hub=5)
    rng = np.random.default_rng(1)
    codes = rng.integers(0, S, n)
    codes[:S] = np.arange(S)
    g = DeviceGraph(A)
    vals, diag = g.scaled(1, torch.float32)
    ld = (S + 7) // 8 * 8
    cur = torch.full((n, ld), 7.0, dtype=torch.float32, device="cuda")
    nxt = torch.zeros_like(cur)
    code_d = torch.as_tensor(codes, dtype=torch.int32, device="cuda")
    _lib.diffuse_onehot(g.indptr, g.indices, vals, diag, code_d, S, cur)
    onehot = np.zeros((n, S))
    onehot[np.arange(n), codes] = 1
    ref = list(orc.diffuse_stepwise(A, onehot, maxnsteps=3))
    got = cur.cpu().numpy()
    assert (got[:, S:] == 0).all()
    np.testing.assert_allclose(got[:, :S], ref[0], rtol=2e-6, atol=1e-9)
    for t in (1, 2):
        _lib.diffuse_step(g.indptr, g.indices, vals, diag, cur, nxt, S)
        cur, nxt = nxt, cur
        np.testing.assert_allclose(cur.cpu().numpy()[:, :S], ref[t], rtol=5e-6, atol=1e-9)
    # per-sample mass is conserved by the column-stochastic operator
    np.testing.assert_allclose(cur.double().sum(0).cpu().numpy()[:S], np.bincount(codes, minlength=S), rtol=1e-5)
    # row kurtosis (auto-stop statistic)
    import scipy.stats as st
    C = np.bincount(codes, minlength=S).astype(float)
    kurt = torch.empty(n, dtype=torch.float64, device="cuda")
    _lib.row_kurtosis(cur, S, torch.as_tensor(1 / C, device="cuda"), kurt)
    x = cur.cpu().numpy()[:, :S].astype(np.float64) / C
    np.testing.assert_allclose(kurt.cpu().numpy(), st.kurtosis(x, axis=1), rtol=1e-9, atol=1e-9)
    # fp64 generic kernel, odd column count
    s0 = rng.normal(size=(n, 5))
    v64, d64 = g.scaled(0.5, torch.float64)
    a = torch.as_tensor(s0, device="cuda")
    b = torch.empty_like(a)
    _lib.diffuse_step(g.indptr, g.indices, v64, d64, a, b, 5)
    np.testing.assert_allclose(b.cpu().numpy(), orc.diffuse(A, s0, 1, self_weight=0.5), rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("n,S,nb", [(3000, 50, 5), (2048, 200, 4), (900, 100, 2), (700, 333, 8)])
def test_fused_qc_step_matches_separate_kernels(cna, n, S, nb):
    """The last diffusion step with the batch-kurtosis epilogue = plain step + cna_batch_kurtosis."""
    import pandas as pd
    import torch
    from cna_b200 import _lib
    from cna_b200.tl import _nam
    from cna_b200.tl._graph import DeviceGraph, _to_dev
    A = _random_graph(n, 9, seed=n + S + nb, hub=5)
    rng = np.random.default_rng(nb)
    codes = rng.integers(0, S, n)
    codes[:S] = np.arange(S)
    g = DeviceGraph(A)
    vals, diag = g.scaled(1, torch.float32)
    ld = (S + 7) // 8 * 8
    cur = torch.zeros((n, ld), dtype=torch.float32, device="cuda")
    _lib.diffuse_onehot(g.indptr, g.indices, vals, diag, torch.as_tensor(codes, dtype=torch.int32, device="cuda"), S, cur)
    labels = pd.Index(np.arange(S))
    batches = pd.Series(rng.permutation(np.arange(S) % nb), index=labels)
    col_batch, batch_inv = _nam._qc_plan(batches, labels, ld, cur.device)
    counts = np.bincount(codes, minlength=S).astype(np.float64)
    inv = torch.as_tensor(1 / counts, device="cuda")
    inv_ld = torch.zeros(ld, dtype=torch.float64, device="cuda")
    inv_ld[:S] = inv
    plain = torch.zeros_like(cur)
    _lib.diffuse_step(g.indptr, g.indices, vals, diag, cur, plain, S)
    ub, order, off = _nam._batch_segments(batches.to_numpy())
    want = torch.empty(n, dtype=torch.float64, device="cuda")
    _lib.batch_kurtosis(plain, inv, _to_dev(order), _to_dev(off), want)
    out = torch.zeros_like(cur)
    kurt = torch.full((n,), -7.0, dtype=torch.float64, device="cuda")
    _lib.diffuse_step_qc(g.indptr, g.indices, vals, diag, cur, out, S, col_batch, inv_ld, batch_inv, kurt)
    assert torch.equal(out[:, :S], plain[:, :S])
    np.testing.assert_allclose(kurt.cpu().numpy(), want.cpu().numpy(), rtol=1e-9, atol=1e-9, equal_nan=True)


@pytest.mark.parametrize("N,n,S,r,nb", [(5000, 50, 50, 6, 5), (3000, 37, 45, 0, 1), (4100, 200, 200, 5, 4),
                                       (1000, 100, 120, 12, 10), (2000, 330, 400, 3, 2)])
def test_resid_pass_gram_null_vs_oracle(cna, N, n, S, r, nb):
    import torch
    from cna_b200 import _lib
    from cna_b200.tl import _nam, _stats
    from oracle import cna_oracle as orc
    rng = np.random.default_rng(N + n)
    counts = rng.integers(20, 60, S).astype(float)
    raw = rng.gamma(2.0, 1.0, (N, S)) * counts  # un-normalised state
    raw[5] = 0.0                                  # zero-variance row
    raw32 = raw.astype(np.float32)
    ld = (S + 7) // 8 * 8
    s = torch.zeros((N, ld), dtype=torch.float32, device="cuda")
    s[:, :S] = torch.as_tensor(raw32)
    colmap = rng.permutation(S)[:n].astype(np.int32)
    batches = rng.integers(0, nb, n) if nb > 1 else np.ones(n)
    if nb > 1:
        batches[:nb] = np.arange(nb)
    ncov = r - (nb if nb > 1 else 0)
    covs = rng.normal(size=(n, ncov)) if ncov > 0 else None
    y = rng.normal(size=n)
    y = (y - y.mean()) / y.std()
    st = _nam.NamState(s, S, pd.Index(np.arange(S)), counts, pd.RangeIndex(N))
    keep = rng.random(N) > 0.02
    st.keep = torch.as_tensor(keep.astype(np.uint8), device="cuda")
    res = _nam.resid_nam_device(st, colmap, covs, batches, y, ridges=[1.0, 0.0] if nb > 1 else None)
    # oracle on the same fp32-rounded input
    X0 = (raw32.astype(np.float64) / counts)[:, colmap]
    valid = keep & (X0.std(axis=1, ddof=1) != 0)
    np.testing.assert_array_equal(res.valid.cpu().numpy().astype(bool), valid)
    o = orc.resid_nam(X0[valid], covs, batches, ridges=[1.0, 0.0] if nb > 1 else None)
    assert o.r == res.r
    np.testing.assert_allclose(res.M, o.M, atol=1e-12)
    xg = res.x.cpu().numpy()
    assert (xg[~valid] == 0).all() and (xg[:, n:] == 0).all()
    np.testing.assert_allclose(xg[valid][:, :n], o.X, rtol=1e-6, atol=2e-6)
    nc = (o.X * y).sum(axis=1) / n
    np.testing.assert_allclose(res.ncorr.cpu().numpy()[valid], nc, rtol=1e-6, atol=1e-7)
    # Gram: tensor-core (if enabled) and SIMT kernels against float64
    x64 = xg[:, :n].astype(np.float64)
    Gref = x64.T @ x64
    for simt in (False, True):
        G = torch.zeros((n, n), dtype=torch.float64, device="cuda")
        _lib.gram(res.x, n, G, simt=simt)
        np.testing.assert_allclose(G.cpu().numpy(), Gref, rtol=0, atol=5e-6 * np.abs(Gref).max())
    # permutation engine against the oracle's matrix form
    U, svs, _ = _nam.gram_svd(res.x, n)
    K = 257
    np.random.seed(3)
    bix = _stats.conditional_permutation_indices(batches, K)
    ks = [1, 3, 4] if n < 100 else list(_default_ks(n))
    kmax = max(ks)
    dev = "cuda"
    Kl = 100
    ycond = torch.zeros((res.x.shape[1], 104), dtype=torch.float32, device=dev)
    ssered = torch.empty(K, dtype=torch.float64, device=dev)
    ssefull = torch.empty((K, len(ks)), dtype=torch.float64, device=dev)
    td = lambda a, dt=None: torch.as_tensor(np.ascontiguousarray(a), device=dev, dtype=dt)  # noqa: E731
    _lib.perm_stats(td(y), td(bix.T, torch.int32), td(res.C) if res.r else None,
                    td(res.W_last) if res.r else None, td(U[:, :kmax].T), td(ks, torch.int32),
                    ssered, ssefull, ycond, Kl)
    Z = y[bix]
    Zc = res.M.dot(Z)
    Zc = Zc / Zc.std(axis=0, ddof=1)
    np.testing.assert_allclose(ssered.cpu().numpy(), (Zc * Zc).sum(0), rtol=1e-12)
    for a, k in enumerate(ks):
        sse = ((U[:, :k].dot(U[:, :k].T.dot(Zc)) - Zc) ** 2).sum(0)
        np.testing.assert_allclose(ssefull.cpu().numpy()[:, a], sse, rtol=1e-9)
    yc = ycond.cpu().numpy()
    np.testing.assert_allclose(yc[:n, :Kl], Zc[:, :Kl], rtol=1e-6, atol=1e-7)
    assert (yc[n:] == 0).all() and (yc[:, Kl:] == 0).all()
    # null histogram: exact counts against the same fp32 inputs evaluated in float64, away from
    # rounding distance of an edge
    z = np.abs(x64 @ yc[:n, :Kl].astype(np.float64) / n)
    mx = max(np.abs(res.ncorr.cpu().numpy()).max(), 0.001)
    thr = np.arange(mx / 4, mx, mx / 400)
    edges = _stats.threshold_edges(thr)
    hist = torch.zeros((Kl, len(thr)), dtype=torch.int32, device=dev)
    _lib.null_hist(res.x, n, ycond, Kl, td(edges), float(edges[0]), hist)
    tails = _stats.tails_from_hist(hist.cpu().numpy().astype(np.int64))
    z2 = z ** 2
    lo = (z2[:, :, None] >= edges[None, None, :] * (1 + 1e-5)).sum(0)
    hi = (z2[:, :, None] >= edges[None, None, :] * (1 - 1e-5)).sum(0)
    assert (tails >= lo).all() and (tails <= hi).all()
    assert tails.sum() > 0
    # observed histograms + per-cell lookup: exact integer logic
    ncorr = res.ncorr.cpu().numpy()
    obs = torch.zeros((2, len(thr)), dtype=torch.int32, device=dev)
    _lib.obs_hist(res.ncorr, res.valid, td(edges), td(thr), obs[0], obs[1])
    oh = obs.cpu().numpy().astype(np.int64)
    np.testing.assert_array_equal(_stats.tails_from_hist(oh[0]), orc.tail_counts(thr, ncorr[valid])[0])
    np.testing.assert_array_equal(_stats.tails_from_hist(oh[1]),
                                  [(np.abs(ncorr[valid]) > t).sum() for t in thr])
    fdr = rng.random(len(thr))
    pmin = np.fmin.accumulate(fdr)
    coef = torch.empty(N, dtype=torch.float64, device=dev)
    cf = torch.empty(N, dtype=torch.float64, device=dev)
    _lib.cell_fdr(res.ncorr, res.valid, td(thr), td(pmin), coef, cf)
    full = np.where(valid, ncorr, np.nan)
    want = orc.cell_fdr_lookup(full, pd.DataFrame({"threshold": thr, "fdr": fdr}))
    np.testing.assert_array_equal(cf.cpu().numpy(), want)
    np.testing.assert_array_equal(np.isnan(coef.cpu().numpy()), ~valid)


def test_cell_reordering_is_transparent(cna, demo, synth, monkeypatch):
    """The Cuthill-McKee cell order (tl/_graph.py, csrc/reorder.cu) only renames cells on the device:
    forced on for the small golden graphs, every output must still match the reference, and the
    diffusion must agree with the run in the original order to rounding of the edge normalisation
    (rows keep their edges in the original order, so the sums themselves are performed identically)."""
    import torch
    from cna_b200.tl._graph import DeviceGraph
    A = cases.demo_anndata().obsp["connectivities"]
    gr = DeviceGraph(A, reorder=True)
    N = A.shape[0]
    order, inv = gr.order.cpu().numpy(), gr.inv.cpu().numpy()
    assert sorted(order.tolist()) == list(range(N)) and (inv[order] == np.arange(N)).all()
    # locality: the reordered graph is banded
    Ap = sp.csr_matrix((gr.data.cpu().numpy(), gr.indices.cpu().numpy(), gr.indptr.cpu().numpy()), shape=A.shape)
    assert abs(Ap - A[order][:, order]).max() == 0
    rows = np.repeat(np.arange(N), np.diff(Ap.indptr))
    assert np.mean(np.abs(rows - Ap.indices)) < 0.5 * np.mean(np.abs(np.repeat(np.arange(N), np.diff(A.indptr)) - A.indices))

    class D:
        pass
    rng = np.random.default_rng(0)
    s0 = rng.normal(size=(N, 3))
    d = D()
    d.obsp = {"connectivities": A}
    monkeypatch.setenv("CNA_B200_REORDER", "0")
    plain = cna.tl.diffuse(d, s0, 2)
    monkeypatch.setenv("CNA_B200_REORDER", "1")
    reord = cna.tl.diffuse(d, s0, 2)
    # same sums in the same order; only the column sums (fp64 atomics) differ in the last bits
    np.testing.assert_allclose(plain, reord, rtol=1e-11, atol=1e-15)
    for name in list(cases.DEMO_CASES)[:3]:
        arrays, scalars = demo
        spec = cases.DEMO_CASES[name]
        data, kwargs = cases.build_demo_case(cases.load_demo_graph(), spec)
        res, warns = helpers.run_association(cna.tl.association, data, kwargs, spec.get("np_seed"))
        helpers.assert_matches_golden(res, data, kwargs.get("key_added", "coef"), arrays, scalars, name,
                                      warns=warns, rtol=RTOL, atol=1e-9, fp32=True)
    for name in list(cases.SYNTH_CASES):
        arrays, scalars = synth
        spec = cases.SYNTH_CASES[name]
        data, kwargs = cases.build_synth_case(spec, helpers.synth_raw(arrays, name))
        res, warns = helpers.run_association(cna.tl.association, data, kwargs)
        helpers.assert_matches_golden(res, data, "coef", arrays, scalars, name, warns=warns, rtol=RTOL,
                                      atol=1e-9, fp32=True)
    # nam() / svd_nam on the reordered graph: labels and keep mask in the caller's order
    ref_spec = dict(y="case", covs=["male"], batches="batch")
    data, kw = cases.build_demo_case(cases.load_demo_graph(), ref_spec)
    monkeypatch.setenv("CNA_B200_REORDER", "1")
    nam_r, keep_r = cna.tl.nam(data, "id", batches=kw["batches"], nsteps=3)
    monkeypatch.setenv("CNA_B200_REORDER", "0")
    nam_p, keep_p = cna.tl.nam(data, "id", batches=kw["batches"], nsteps=3)
    np.testing.assert_array_equal(keep_r, keep_p)
    assert (nam_r.columns == nam_p.columns).all()
    np.testing.assert_allclose(nam_r.to_numpy(), nam_p.to_numpy(), rtol=1e-6, atol=1e-12)


def test_obs_columns_do_not_alias_the_staging_buffer(cna):
    """The per-cell results travel through a reusable pinned staging buffer: a second call must not
    change the columns written by the first."""
    data, kw = cases.build_demo_case(cases.load_demo_graph(), dict(y="case", batches="batch", nsteps=2,
                                                                    Nnull=50, seed=1))
    cna.tl.association(data, key_added="first", **{k: v for k, v in kw.items() if k != "key_added"})
    first = data.obs["first"].to_numpy().copy()
    kw2 = dict(kw)
    kw2["y"] = -kw["y"]
    cna.tl.association(data, key_added="second", **{k: v for k, v in kw2.items() if k != "key_added"})
    np.testing.assert_array_equal(data.obs["first"].to_numpy(), first)
    np.testing.assert_allclose(data.obs["second"].to_numpy(), -first, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("n,S,deg", [(3000, 50, 7), (5000, 200, 30), (700, 333, 9), (257, 3, 4)])
def test_tiled_diffusion_step_is_bit_identical(cna, n, S, deg):
    """cna_diffuse_step_f32_tiled (shared-memory-staged rows: TMA tile::gather4 or cp.async) performs the
    additions of every row in the same order as cna_diffuse_step_f32: identical bits, hub row included."""
    import torch
    from cna_b200 import _lib
    from cna_b200.tl import _graph
    A = _random_graph(n, deg, seed=n + S, hub=n // 2 if n < 2000 else None)  # hub row: n / 3 neighbours
    g = _graph.DeviceGraph(A, reorder=False)
    vals, diag = g.scaled(1, torch.float32)
    ld = (S + 7) // 8 * 8
    src = torch.zeros((n, ld), device="cuda")
    src[:, :S] = torch.rand((n, S), device="cuda")
    ref = torch.empty_like(src)
    _lib.diffuse_step(g.indptr, g.indices, vals, diag, src, ref, S)
    plan = _graph.TilePlan(g.indptr, g.indices, vals, g.n)
    for mode in (0, 1):
        out = torch.full_like(src, float("nan"))
        _lib.diffuse_step_tiled(g.indptr, plan, diag, src, out, S, stage_mode=mode)
        assert torch.equal(out[:, :S], ref[:, :S])


def test_device_median_numpy_semantics(cna):
    """cna_median_f64 (radix select, result left on the device) against np.median: odd / even sizes,
    duplicates around the middle, masks, the skip pattern, infinities, signed zeros, NaN, empty."""
    import torch
    from cna_b200 import _lib
    from cna_b200.tl._nam import device_median, median_device
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 7, 10, 1001, 4096, 100_003, 1_000_000):
        v = rng.normal(size=n)
        assert device_median(torch.as_tensor(v).cuda()) == np.median(v)
        w = np.round(v, 1 if n > 100 else 0)  # many duplicates, also across the middle
        assert device_median(torch.as_tensor(w).cuda()) == np.median(w)
        k = rng.gamma(2.0, 1.0, n) * 10.0 ** rng.integers(-3, 3)  # clustered exponents, like kurtoses
        assert device_median(torch.as_tensor(k).cuda()) == np.median(k)
        mask = rng.random(n) > 0.4
        mask[0] = True
        got = median_device(torch.as_tensor(v).cuda(), valid=torch.as_tensor(mask).cuda()).cpu().numpy()
        assert got[0] == np.median(v[mask]) and got[1] == mask.sum()
    v = rng.normal(size=1000)
    v[::7] = np.inf
    v[1::7] = -np.inf
    v[2::50] = 0.0
    v[3::50] = -0.0
    assert device_median(torch.as_tensor(v).cuda()) == np.median(v)
    skip = np.full(300, _lib.median_skip_value()[0])
    both = torch.as_tensor(np.concatenate([v, skip])).cuda()
    got = median_device(both).cpu().numpy()
    assert got[0] == np.median(v) and got[1] == len(v)
    v[3] = np.nan
    assert np.isnan(device_median(torch.as_tensor(v).cuda()))
    masked = np.ones(len(v), dtype=bool)
    masked[3] = False  # a masked NaN does not count
    assert device_median(torch.as_tensor(v).cuda(), valid=torch.as_tensor(masked).cuda()) == np.median(v[masked])
    assert np.isnan(device_median(torch.empty(0, dtype=torch.float64, device="cuda")))
    assert np.isnan(device_median(torch.ones(5, dtype=torch.float64, device="cuda"),
                                  valid=torch.zeros(5, dtype=torch.uint8, device="cuda")))


def test_device_fdr_thresholds_match_numpy(cna):
    """cna_fdr_thresholds against np.arange(m/4, m, m/400) and _stats.threshold_edges, bit for bit."""
    import torch
    from cna_b200 import _lib
    from cna_b200.tl import _stats
    from cna_b200.tl._association import THRESHOLD_CAP
    rng = np.random.default_rng(1)
    ms = np.concatenate([10 ** rng.uniform(-4, 0.3, 400), [0.0, 1e-9, 0.001, 0.25, 1.0]])
    thr_d = torch.empty(THRESHOLD_CAP, dtype=torch.float64, device="cuda")
    edges_d = torch.empty(THRESHOLD_CAP, dtype=torch.float64, device="cuda")
    cnt_d = torch.empty(1, dtype=torch.int32, device="cuda")
    for m in ms:
        _lib.fdr_thresholds(torch.tensor([m], dtype=torch.float64, device="cuda"), thr_d, edges_d, cnt_d)
        mm = max(float(m), 0.001)
        want = np.arange(mm / 4, mm, mm / 400)
        T = int(cnt_d.item())
        assert T == len(want)
        np.testing.assert_array_equal(thr_d.cpu().numpy()[:T], want)
        np.testing.assert_array_equal(edges_d.cpu().numpy()[:T], _stats.threshold_edges(want))
        assert (thr_d.cpu().numpy()[T:] == 0).all()


def _default_ks(n):
    from cna_b200.tl._association import default_ks
    return default_ks(n)


def test_batch_kurtosis_qc_vs_oracle(cna):
    import torch
    from cna_b200 import _lib
    from cna_b200.tl import _nam
    from oracle import cna_oracle as orc
    rng = np.random.default_rng(11)
    N, S = 4000, 60
    counts = rng.integers(20, 60, S).astype(float)
    raw32 = (rng.gamma(2.0, 1.0, (N, S)) * counts).astype(np.float32)
    raw32[:200, :10] *= 30  # batch-specific neighbourhoods
    s = torch.zeros((N, 64), dtype=torch.float32, device="cuda")
    s[:, :S] = torch.as_tensor(raw32)
    for nb in (2, 6, 37):
        b = pd.Series(rng.integers(0, nb, S), index=np.arange(S))
        b.iloc[:nb] = np.arange(nb)
        st = _nam.NamState(s, S, pd.Index(np.arange(S)), counts, pd.RangeIndex(N))
        _nam._qc_device(st, b)
        X = raw32.astype(np.float64) / counts
        keep, thr = orc.qc_keep(X, b.to_numpy())
        np.testing.assert_array_equal(_nam.keep_mask(st), keep)
        assert st.qc_threshold == pytest.approx(thr, rel=1e-12)


def test_large_size_properties(cna):
    """Size-independent properties at a benchmark-like shape (100k cells): mass conservation and
    row-stochasticity of the NAM, Gram trace = N'(n-1), idempotent re-run, resident == host path."""
    import torch
    from cna_b200 import synth
    from cna_b200.tl import _nam
    data, meta = synth.make_dataset(100_000, 100, 15, seed=0)
    st = _nam._nam_device(data, "id", nsteps=3)
    x = st.s[:, :100].double() * st.inv_count
    np.testing.assert_allclose(x.sum(0).cpu().numpy(), 1.0, rtol=1e-5)      # each sample's NAM row sums to 1
    np.testing.assert_allclose(st.s[:, :100].double().sum(0).cpu().numpy(), st.counts, rtol=1e-5)
    kw = dict(y=meta.case, sid_name="id", batches=meta.batch, covs=meta[["age"]], nsteps=3, Nnull=1000, seed=0)
    import warnings
    warnings.simplefilter("ignore")
    r1 = cna.tl.association(data, return_full=True, **kw)
    c1, f1 = data.obs["coef"].to_numpy().copy(), data.obs["coef_fdr"].to_numpy().copy()
    nk = int(r1.kept.sum())
    n = len(meta)
    assert r1.namresid_svs.shape[0] == 15
    np.testing.assert_allclose(r1.namresid_varexp.sum() * n * nk, nk * (n - 1), rtol=1e-5)
    h = cna.tl.to_device(data)
    p2 = cna.tl.association(h, **kw)
    assert p2 == r1.p
    np.testing.assert_array_equal(data.obs["coef"].to_numpy(), c1)
    np.testing.assert_array_equal(data.obs["coef_fdr"].to_numpy(), f1)
    assert np.all(np.diff(r1.fdrs.num_detected.to_numpy()) <= 0)               # tail counts are sorted
    assert ((f1 >= 0) & (f1 <= 1) | np.isnan(f1)).all()


@pytest.mark.parametrize("n,r,ks", [(200, 5, [4, 8, 12, 16]), (50, 6, [1, 2, 3, 4]), (37, 0, [1, 3, 4]),
                                    (500, 12, [10, 20, 30, 40]), (12, 1, [1])])
def test_device_f_survival_matches_scipy(cna, n, r, ks):
    """cna_perm_minp: F survival function (incomplete beta continued fraction, fp64) and the min over
    ks, against scipy.stats.f.sf / nanargmin (_association.py:45-46, :53-60)."""
    import scipy.stats as st
    import torch
    from cna_b200 import _lib
    rng = np.random.default_rng(n + r)
    K = 4000
    ssered = rng.uniform(0.5 * n, 1.5 * n, K)
    # r2 from tiny to ~1 (log-uniform in 1 - r2 as well), non-increasing SSE along ks
    frac = np.sort(rng.uniform(0, 1, (K, len(ks))), axis=1)[:, ::-1] ** rng.integers(1, 12, (K, 1))
    ssefull = ssered[:, None] * np.clip(frac, 1e-12, 1.0)
    ssefull[0] = ssered[0]                 # f = 0  -> p = 1
    ssefull[1, -1] = 0.0                   # f = inf -> p = 0
    ksa = np.asarray(ks, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        f = ((ssered[:, None] - ssefull) / ksa) / (ssefull / n)
        want = st.f.sf(f, ksa, n - (1 + r + ksa))
    dev = "cuda"
    minp = torch.empty(K, dtype=torch.float64, device=dev)
    argk = torch.empty(K, dtype=torch.int32, device=dev)
    r2 = torch.empty(K, dtype=torch.float64, device=dev)
    _lib.perm_minp(torch.as_tensor(ssered, device=dev), torch.as_tensor(np.ascontiguousarray(ssefull), device=dev),
                   torch.as_tensor(np.asarray(ks, dtype=np.int32), device=dev), n, r, minp, argk, r2)
    got, ga = minp.cpu().numpy(), argk.cpu().numpy()
    wmin = np.nanmin(want, axis=1)
    # (below ~1e-280 scipy's own value degrades into the denormal range; nothing is decided there)
    np.testing.assert_allclose(got, wmin, rtol=1e-10, atol=1e-250)
    # the chosen k agrees wherever the two smallest p-values are not within rounding of each other
    srt = np.sort(want, axis=1)
    clear = (len(ks) == 1) | ((srt[:, min(1, len(ks) - 1)] > srt[:, 0] * (1 + 1e-9)) & (srt[:, 0] > 1e-250))
    np.testing.assert_array_equal(ga[clear], np.nanargmin(want, axis=1)[clear])
    rows = np.arange(K)
    np.testing.assert_allclose(r2.cpu().numpy(), 1 - ssefull[rows, ga] / ssered, rtol=1e-14)


@pytest.mark.parametrize("N,S,k,nsteps", [(30_000, 500, 15, 3), (40_000, 64, 10, None)])
def test_end_to_end_vs_oracle_at_other_shapes(cna, N, S, k, nsteps):
    """Whole association() against the oracle on seeded synthetic data at shapes the golden fixtures
    do not cover: 500 samples (config E's sample count: n > 256 takes the CUDA-core Gram, the
    null GEMM runs 32 k-steps, 16 column groups in the row pass) and the kurtosis auto-stop rule
    (nsteps=None) on a graph large enough to be stored in the reordered cell order."""
    import warnings
    from cna_b200 import synth
    from oracle import cna_oracle as orc
    data, meta = synth.make_dataset(N, S, k, seed=3, dim=8)
    kw = dict(y=meta.case, sid_name="id", batches=meta.batch, covs=meta[["age"]], nsteps=nsteps, Nnull=300, seed=4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        d_cpu = type(data)(data.obs.copy(), data.obsp["connectivities"])
        want = orc.association(d_cpu, return_full=True, **kw)
        import os
        old = os.environ.get("CNA_B200_REORDER")
        os.environ["CNA_B200_REORDER"] = "1"
        try:
            d_gpu = type(data)(data.obs.copy(), data.obsp["connectivities"])
            got = cna.tl.association(d_gpu, return_full=True, **kw)
            d_gpu2 = type(data)(data.obs.copy(), data.obsp["connectivities"])
            p_fast = cna.tl.association(d_gpu2, **kw)  # device F survival + leading eigenpairs only
        finally:
            if old is None:
                os.environ.pop("CNA_B200_REORDER", None)
            else:
                os.environ["CNA_B200_REORDER"] = old
    assert got.p == want.p == p_fast
    assert int(got.k) == int(want.k) and list(got.ks) == list(want.ks) and got.r == want.r
    np.testing.assert_array_equal(got.kept, want.kept)
    np.testing.assert_allclose(got.ncorrs.to_numpy(), want.ncorrs.to_numpy(), rtol=RTOL, atol=1e-6)
    npc = min(len(got.namresid_svs), 10)
    np.testing.assert_allclose(got.namresid_svs.to_numpy()[:npc], want.namresid_svs.to_numpy()[:npc], rtol=RTOL)
    np.testing.assert_allclose(got.nullminps, want.nullminps, rtol=1e-4, atol=1e-12)
    np.testing.assert_allclose(got.r2, want.r2, rtol=RTOL)
    a, b = d_gpu.obs["coef_fdr"].to_numpy(), d_cpu.obs["coef_fdr"].to_numpy()
    c = d_gpu2.obs["coef_fdr"].to_numpy()
    np.testing.assert_array_equal(a, c)
    np.testing.assert_array_equal(d_gpu.obs["coef"].to_numpy(), d_gpu2.obs["coef"].to_numpy())
    # FDR-passing sets agree except for cells whose |coefficient| is within rounding of the threshold
    for level, thr in ((0.05, want.fdr_5p_t), (0.1, want.fdr_10p_t)):
        if thr is None:
            continue
        edge = np.abs(np.abs(np.nan_to_num(d_cpu.obs["coef"].to_numpy())) - thr) < 1e-5 * max(thr, 1e-3)
        assert ((a <= level) == (b <= level))[~edge].all()


def test_config_B_vs_oracle(cna):
    """BASELINE.json configs[1] at its exact shape (100 000 cells, 100 samples, k = 15, s = 3, 1000
    permutations; bench.py --config B builds the same dataset): whole association() against the
    oracle.  Exact: p, k, the kept cells; the FDR-passing sets agree except for cells within rounding of
    the threshold; singular values and coefficients to 1e-5."""
    import warnings
    from cna_b200 import synth
    from oracle import cna_oracle as orc
    data, meta = synth.make_dataset(100_000, 100, 15, seed=0)
    kw = dict(y=meta.case, sid_name="id", batches=meta.batch, covs=meta[["age"]], nsteps=3, Nnull=1000, seed=0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        d_cpu = type(data)(data.obs.copy(), data.obsp["connectivities"])
        want = orc.association(d_cpu, return_full=True, **kw)
        d_gpu = type(data)(data.obs.copy(), data.obsp["connectivities"])
        got = cna.tl.association(d_gpu, return_full=True, **kw)
        d_res = type(data)(data.obs.copy(), data.obsp["connectivities"])
        p_res = cna.tl.association(cna.tl.to_device(d_res), **kw)  # the path bench.py's `value` times
    assert got.p == want.p == p_res
    assert int(got.k) == int(want.k) and list(got.ks) == list(want.ks) and got.r == want.r
    np.testing.assert_array_equal(got.kept, want.kept)
    np.testing.assert_allclose(got.ncorrs.to_numpy(), want.ncorrs.to_numpy(), rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(got.namresid_svs.to_numpy()[:10], want.namresid_svs.to_numpy()[:10], rtol=RTOL)
    np.testing.assert_allclose(got.nullminps, want.nullminps, rtol=1e-4, atol=1e-12)
    np.testing.assert_allclose(got.fdrs.threshold.to_numpy(), want.fdrs.threshold.to_numpy(), rtol=RTOL)
    a, b = d_gpu.obs["coef_fdr"].to_numpy(), d_cpu.obs["coef_fdr"].to_numpy()
    np.testing.assert_array_equal(d_res.obs["coef_fdr"].to_numpy(), a)
    np.testing.assert_array_equal(d_res.obs["coef"].to_numpy(), d_gpu.obs["coef"].to_numpy())
    for level, thr in ((0.05, want.fdr_5p_t), (0.1, want.fdr_10p_t)):
        if thr is None:
            assert (a <= level).sum() == 0
            continue
        edge = np.abs(np.abs(np.nan_to_num(d_cpu.obs["coef"].to_numpy())) - thr) < 1e-5 * max(thr, 1e-3)
        assert ((a <= level) == (b <= level))[~edge].all()
        assert (b <= level).sum() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("n,nb,num,pre", [(200, 4, 2000, 0), (50, 1, 333, 1), (333, 7, 257, 0), (12, 12, 40, 3),
                                          (201, 3, 1001, 0), (7, 1, 1, 0), (64, 2, 10000, 1)])
def test_device_permutation_draw_is_bit_exact(cna, n, nb, num, pre):
    """cna_perm_draw_device against the reference's numpy call sequence (_stats.py:8-16 and :20-32): the
    same index matrices bit for bit and the same generator state left behind (key, position, cached
    deviate), for even / odd deviate counts and a cached deviate pending on entry."""
    import torch
    from cna_b200.tl import _stats
    rng = np.random.default_rng(n + nb)
    B = rng.integers(0, nb, n)
    B[:nb] = np.arange(nb)
    y = rng.normal(size=n)
    np.random.seed(n)
    np.random.randn(pre)
    want, after_w, state_w = _stats.conditional_permutation_indices(B, num), np.random.randn(5), np.random.get_state()
    np.random.seed(n)
    np.random.randn(pre)
    draw = _stats.PermutationDraw(y, B, None, num, device=torch.device("cuda", 0), engine="device")
    got_d = draw.result_device(torch.device("cuda", 0))
    torch.cuda.synchronize()
    got = draw.result()
    after_g, state_g = np.random.randn(5), np.random.get_state()
    assert got.dtype == np.int32 and got.shape == (num, n)
    np.testing.assert_array_equal(got, want.T)
    np.testing.assert_array_equal(got_d.cpu().numpy(), want.T)
    np.testing.assert_array_equal(after_g, after_w)
    np.testing.assert_array_equal(state_g[1], state_w[1])
    assert state_g[2:] == state_w[2:]
    # donor-level permutations (_stats.py:20-32)
    G = np.arange(n) // 2
    Y = (G % 3 == 0).astype(float)
    np.random.seed(n + 1)
    want, after_w = _stats.grouplevel_permutation_indices(G, Y, num), np.random.randn(3)
    np.random.seed(n + 1)
    draw = _stats.PermutationDraw(Y, None, G, num, device=torch.device("cuda", 0), engine="device")
    got = draw.result_device(torch.device("cuda", 0)).cpu().numpy()
    draw.cancel()
    np.testing.assert_array_equal(got, want.T)
    np.testing.assert_array_equal(np.random.randn(3), after_w)


@pytest.mark.gpu
def test_device_draw_and_host_draw_give_the_same_association(cna):
    """The whole call with the device engine (default) and with the host engine: identical p, k, kept set
    and per-cell columns, and the same generator state afterwards."""
    spec = dict(y="case", covs=["male"], batches="batch", seed=3, Nnull=500, nsteps=3)
    out = []
    for host in (False, True):
        d, kw = cases.build_demo_case(cases.load_demo_graph(), spec)
        if host:
            kw["_host_draw"] = True
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            p = cna.tl.association(d, **kw)
        out.append((p, d.obs["coef"].to_numpy().copy(), d.obs["coef_fdr"].to_numpy().copy(), np.random.randn(3)))
    assert out[0][0] == out[1][0]
    np.testing.assert_array_equal(out[0][1], out[1][1])
    np.testing.assert_array_equal(out[0][2], out[1][2])
    np.testing.assert_array_equal(out[0][3], out[1][3])


@pytest.mark.gpu
@pytest.mark.parametrize("n,r,kmax,K", [(700, 5, 56, 300), (1000, 3, 80, 200), (64, 0, 12, 500)])
def test_perm_stats_many_samples_streams_the_pcs(cna, n, r, kmax, K):
    """cna_perm_stats beyond the shared-memory staging limit (kmax * n doubles > 200 KB from n ~ 570 with
    the default ks): the PC matrix is then read where it lies.  Against the reference's arithmetic
    (_association.py:35-52) in numpy."""
    import torch
    from cna_b200 import _lib
    rng = np.random.default_rng(n)
    y = rng.normal(size=n)
    perm = np.stack([rng.permutation(n) for _ in range(K)]).astype(np.int32)
    U, _ = np.linalg.qr(rng.normal(size=(n, kmax)))
    ks = np.unique(np.linspace(max(1, kmax // 4), kmax, 4).astype(np.int32))
    if r:
        C = rng.normal(size=(n, r))
        C -= C.mean(axis=0)
        W = np.linalg.solve(C.T @ C + 1e-3 * np.eye(r), C.T)
        M = np.eye(n) - C @ W
    else:
        C = W = None
        M = np.eye(n)
    td = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), device="cuda", dtype=dt)  # noqa: E731
    ssered = torch.empty(K, dtype=torch.float64, device="cuda")
    ssefull = torch.full((K, len(ks)), float("nan"), dtype=torch.float64, device="cuda")
    _lib.perm_stats(td(y), td(perm, torch.int32), td(C) if r else None, td(W) if r else None, td(U.T),
                    td(ks, torch.int32), ssered, ssefull, None, 0)
    z = M @ y[perm].T                      # n x K
    z = z / z.std(axis=0, ddof=1)
    want_red = (z * z).sum(axis=0)
    want_full = np.stack([((z - U[:, :k] @ (U[:, :k].T @ z)) ** 2).sum(axis=0) for k in ks], axis=1)
    np.testing.assert_allclose(ssered.cpu().numpy(), want_red, rtol=1e-12)
    np.testing.assert_allclose(ssefull.cpu().numpy(), want_full, rtol=1e-9)


@pytest.mark.gpu
def test_ks_in_any_order_and_with_duplicates(cna):
    """The reference accepts ks in any order (_association.py:25-61 loops over them as given); the device
    kernels want them ascending and distinct, so the host sorts, de-duplicates and maps back: same p, k and
    per-k statistics as the sorted call, and as the oracle."""
    from oracle import cna_oracle as orc
    spec = dict(y="case", covs=["male"], batches="batch", seed=2, Nnull=200, nsteps=3)
    out = {}
    for name, ks in (("sorted", [2, 4, 6, 8]), ("shuffled", [6, 2, 8, 4, 4, 2])):
        d, kw = cases.build_demo_case(cases.load_demo_graph(), dict(spec, ks=ks))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out[name] = (cna.tl.association(d, return_full=True, **kw), d.obs["coef_fdr"].to_numpy().copy())
    a, b = out["sorted"], out["shuffled"]
    assert a[0].p == b[0].p and int(a[0].k) == int(b[0].k)
    np.testing.assert_array_equal(a[0].nullminps, b[0].nullminps)
    np.testing.assert_array_equal(a[1], b[1])
    d, kw = cases.build_demo_case(cases.load_demo_graph(), dict(spec, ks=[6, 2, 8, 4, 4, 2]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = orc.association(d, return_full=True, **kw)
    assert b[0].p == want.p and int(b[0].k) == int(want.k)


@pytest.mark.gpu
def test_staged_host_upload_round_trips(cna):
    """cna_host_upload: pageable buffers staged through the library's page-locked ring (several threads, 4 MB
    chunks, a ring smaller than the buffer so that slots are recycled), page-locked buffers in one copy, the
    background form, odd sizes — the device copy equals the source bit for bit."""
    import torch
    from cna_b200 import _lib
    rng = np.random.default_rng(0)
    for n, dtype in ((30_000_001, np.float64), (5_000_003, np.int32), (17, np.float64), (1 << 20, np.int32)):
        a = rng.integers(-2 ** 31, 2 ** 31 - 1, n).astype(dtype) if dtype == np.int32 else rng.normal(size=n)
        up = _lib.HostUpload(a, torch.device("cuda", 0))
        torch.cuda.synchronize()
        np.testing.assert_array_equal(up.wait().cpu().numpy(), a)
        bg = _lib.HostUpload(a, torch.device("cuda", 0), background=True, n_threads=3)
        t = bg.wait()
        torch.cuda.synchronize()
        np.testing.assert_array_equal(t.cpu().numpy(), a)
    pinned = torch.empty(3_000_000, dtype=torch.float64, pin_memory=True)
    pinned.copy_(torch.as_tensor(rng.normal(size=3_000_000)))
    up = _lib.HostUpload(pinned.numpy(), torch.device("cuda", 0))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(up.wait().cpu().numpy(), pinned.numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["batchy_qc", "all_batchy_ridgewalk"])
def test_qc_decided_after_the_residualisation_pass(cna, synth, name, monkeypatch):
    """When every sample is selected the QC statistic is produced by cna_resid_pass (qc_out) and the decision
    applied by cna_qc_fixup (the route the golden cases above take); with CNA_B200_QC_IN_SPMM=1 the statistic
    comes from the last diffusion step's epilogue and the decision is taken inside the pass.  On the two golden
    cases whose QC drops cells (one of them walks several ridges, so later stages reuse the statistic as an
    input): the same kept set, coefficients, FDRs and p; in the first case cells are indeed dropped."""
    from cna_b200.tl import _association as A_
    arrays, scalars = synth
    spec = cases.SYNTH_CASES[name]
    res = {}
    for mode in ("pass", "spmm"):
        if mode == "spmm":
            monkeypatch.setenv("CNA_B200_QC_IN_SPMM", "1")
        data, kwargs = cases.build_synth_case(spec, helpers.synth_raw(arrays, name))
        if mode == "pass":
            assert A_._all_samples_selected(pd.Index(np.unique(data.obs["id"])), kwargs["y"], kwargs.get("batches"),
                                            kwargs.get("covs"))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            full = cna.tl.association(data, return_full=True, **kwargs)
        res[mode] = (full, data.obs["coef"].to_numpy().copy(), data.obs["coef_fdr"].to_numpy().copy())
    a, b = res["pass"], res["spmm"]
    np.testing.assert_array_equal(a[0].kept, b[0].kept)
    if name == "batchy_qc":
        assert 0 < (~a[0].kept).sum() < len(a[0].kept)
    assert a[0].p == b[0].p and int(a[0].k) == int(b[0].k)
    np.testing.assert_allclose(a[1], b[1], rtol=1e-12, atol=1e-14, equal_nan=True)
    np.testing.assert_allclose(a[2], b[2], rtol=1e-12, atol=1e-14, equal_nan=True)
