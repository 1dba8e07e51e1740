"""CPU tests: the C-ABI library builds / loads / exports what include/cna_b200.h declares, and the
host-side logic of ``cna_b200.tl`` (input checks, permutation drawing, FDR bookkeeping, the small
design algebra) agrees with the oracle.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest
import torch

from oracle import cna_oracle as orc
from tests.golden import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from cna_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "cna_b200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|const char \*|void \*)\s*\*?\s*(cna_\w+)\s*\(", header, flags=re.M))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/cna_b200.h but not exported"
    assert declared == set(_lib.EXPORTS)
    assert lib.cna_abi_version() == 5
    assert isinstance(lib.cna_last_error(), bytes)
    # every entry point is documented: a "replaces:" citation in the header, a row in INTEGRATION.md
    integration = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for name in declared:
        assert name in integration, f"{name} is missing from INTEGRATION.md"
    assert header.count("replaces:") >= 20


def test_no_cpu_fallback():
    """Product code must fail loudly without a CUDA device / with CPU tensors."""
    from cna_b200 import _lib
    from cna_b200.tl import _graph
    with pytest.raises(_lib.CnaError, match="CUDA tensor"):
        _lib.absmax(torch.zeros(3, dtype=torch.float64), None, torch.zeros(1, dtype=torch.float64))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU path"):
            _graph.device()
    src = "".join(open(os.path.join(ROOT, "cna_b200", d, f)).read()
                  for d in ("", "tl") for f in os.listdir(os.path.join(ROOT, "cna_b200", d)) if f.endswith(".py"))
    assert "oracle" not in src.replace("the oracle", ""), "product code must not import the oracle"


def test_default_ks():
    from cna_b200.tl._association import default_ks
    assert list(default_ks(50)) == [1, 2, 3, 4]
    assert list(default_ks(100)) == [2, 4, 6, 8]
    assert list(default_ks(200)) == [4, 8, 12, 16]
    assert list(default_ks(500)) == [10, 20, 30, 40]
    for n in (10, 13, 24, 77, 333):
        assert list(default_ks(n)) == list(orc.default_ks(n))


def test_check_inputs_matches_oracle():
    from cna_b200.tl._association import check_inputs
    data = cases.demo_anndata()
    meta = cases.demo_sample_meta()
    y = meta.case.copy()
    y.iloc[3] = np.nan
    covs = meta[["male"]].copy()
    covs.iloc[7, 0] = np.nan
    extra = pd.concat([y, pd.Series([1.0], index=[99])])
    for args in [(y, None, None, None), (y, meta.batch, covs, None), (extra, None, None, None),
                 (y, None, covs, meta.batch)]:
        a = check_inputs(data, args[0], "id", args[1], args[2], args[3], False)
        b = orc.check_inputs(data, args[0], "id", args[1], args[2], args[3], False)
        pd.testing.assert_series_equal(a[0], b[0])
        np.testing.assert_array_equal(np.asarray(a[1]), np.asarray(b[1]))
    for bad, exc in [((y.to_numpy(), None, None, None), TypeError), ((y, [1], None, None), TypeError),
                     ((y, None, meta.male, None), TypeError), ((y, None, None, [1]), TypeError),
                     ((y.iloc[:30], None, None, None), ValueError), ((y, meta.batch, None, meta.batch), ValueError)]:
        with pytest.raises(exc):
            check_inputs(data, bad[0], "id", bad[1], bad[2], bad[3], False)
    few = y.copy()
    few.iloc[:45] = np.nan
    with pytest.raises(ValueError, match="fewer than 10"):
        check_inputs(data, few, "id", None, None, None, False)
    check_inputs(data, few, "id", None, None, None, True)


def test_permutation_indices_are_bit_exact():
    from cna_b200.tl import _stats
    from cna_b200.tl._stats import PermutationDraw as _PermutationJob
    rng = np.random.default_rng(0)
    B = rng.integers(0, 4, 57)
    Y = rng.normal(size=57)
    np.random.seed(11)
    want = orc.conditional_permutation(B, Y, 300)
    state_after = np.random.get_state()[1].copy()
    np.random.seed(11)
    bix = _stats.conditional_permutation_indices(B, 300)
    np.testing.assert_array_equal(Y[bix], want)
    np.testing.assert_array_equal(np.random.get_state()[1], state_after)  # same RNG consumption
    np.random.seed(11)
    np.testing.assert_array_equal(_PermutationJob(Y, B, None, 300).result(), bix.T)
    # every column permutes within batches only
    assert all((B[bix[:, k]] == B).all() and len(set(bix[:, k])) == 57 for k in range(300))
    G = np.arange(57) // 3
    Yg = (G % 2).astype(float)
    np.random.seed(5)
    want = orc.grouplevel_permutation(G, Yg, 200)
    np.random.seed(5)
    np.testing.assert_array_equal(Yg[_stats.grouplevel_permutation_indices(G, Yg, 200)], want)
    assert _stats.grouplevel_permutation_indices(G, Y, 10) is None
    np.random.seed(5)
    with pytest.raises(TypeError):
        _PermutationJob(Y, None, G, 10).result()


def test_fdr_bookkeeping_matches_oracle():
    from cna_b200.tl import _stats
    rng = np.random.default_rng(3)
    z = rng.normal(0, 0.2, 5000)
    znull = np.abs(rng.normal(0, 0.08, (5000, 37)))
    mx = max(np.abs(z).max(), 0.001)
    t = np.arange(mx / 4, mx, mx / 400)
    edges = _stats.threshold_edges(t)
    np.testing.assert_array_equal(edges, orc.threshold_edges(t))
    bins = np.concatenate([edges, [np.inf]])
    hist = np.array([np.histogram(c ** 2, bins=bins)[0] for c in znull.T])
    rank_hist = np.histogram(z ** 2, bins=bins)[0]
    np.testing.assert_array_equal(_stats.tails_from_hist(hist), orc.tail_counts(t, znull, faithful=True))
    np.testing.assert_allclose(_stats.fdr_from_counts(hist, rank_hist), orc.empirical_fdrs(z, znull, t), rtol=1e-15)


def test_cumulative_projector_algebra():
    """_nam.py:148 applies M = I - C.W cumulatively; resid_nam_device folds the product into one
    rank-r update: prod_k (I - C W_k) = I - C W_cum."""
    from cna_b200.tl import _nam
    rng = np.random.default_rng(1)
    n = 40
    batches = rng.integers(0, 5, n)
    covs = rng.normal(size=(n, 2))
    C, nb = _nam.design_matrix(covs, batches, n)
    Co, nbo = orc.design_matrix(covs, batches)
    np.testing.assert_array_equal(C, Co)
    assert nb == nbo == 5
    Mprod, Wcum = np.eye(n), np.zeros((C.shape[1], n))
    for ridge in (1e3, 10.0, 0.1, 0.0):
        W = _nam.projector(C, nb, ridge)
        np.testing.assert_allclose(W, orc.projector_stage(C, nb, ridge), rtol=1e-13)
        Mprod = (np.eye(n) - C.dot(W)).dot(Mprod)
        Wcum = Wcum + W - W.dot(C).dot(Wcum)
        np.testing.assert_allclose(np.eye(n) - C.dot(Wcum), Mprod, atol=1e-12)


def test_f_statistics_match_oracle():
    from cna_b200.tl._association import _f_pvalues, _pick
    rng = np.random.default_rng(2)
    n, r, K = 60, 3, 50
    U = np.linalg.qr(rng.normal(size=(n, n)))[0]
    M = np.eye(n) - np.outer(U[:, -1], U[:, -1])
    Z = rng.normal(size=(n, K))
    ks = [2, 5, 9]
    kk, pp, rr = orc.minp_stats_matrix(Z, M, U, ks, n, r)
    Zc = M.dot(Z)
    Zc = Zc / Zc.std(axis=0, ddof=1)
    ssered = (Zc * Zc).sum(0)
    ssefull = np.stack([((U[:, :k].dot(U[:, :k].T.dot(Zc)) - Zc) ** 2).sum(0) for k in ks], axis=1)
    p, r2 = _f_pvalues(ssered, ssefull, ks, n, r)
    k2, p2, r22 = _pick(p, r2, ks)
    np.testing.assert_array_equal(k2, kk)
    np.testing.assert_allclose(p2, pp, rtol=1e-12)
    np.testing.assert_allclose(r22, rr, rtol=1e-12)


def test_fdr_threshold_arithmetic_matches_numpy_arange():
    """csrc/select.cu derives the FDR thresholds on the device with numpy's own sequence of float64
    operations (arange: length = ceil((stop - start) / step), values start + i * ((start + step) - start)).
    This pins that restatement against np.arange itself; the kernel is compared with np.arange on the GPU."""
    import math
    rng = np.random.default_rng(0)
    for it in range(20000):
        m = float(10 ** rng.uniform(-3, 0.5))
        ref = np.arange(m / 4, m, m / 400)
        start, step = m / 4, m / 400
        T = math.ceil((m - start) / step)
        delta = (start + step) - start
        t = start + np.arange(T, dtype=np.float64) * delta
        t[0], t[1] = start, start + step
        assert T == len(ref) and np.array_equal(t, ref)


def test_batch_segments():
    from cna_b200.tl._nam import _batch_segments
    b = np.array([2.0, 0.0, 2.0, 1.0, 0.0, 2.0])
    ub, order, off = _batch_segments(b)
    assert list(ub) == [0.0, 1.0, 2.0] and list(off) == [0, 2, 3, 6]
    assert [sorted(order[off[i]:off[i + 1]]) for i in range(3)] == [[1, 4], [3], [0, 2, 5]]


def test_native_legacy_rng_is_bit_exact():
    """csrc/perm_host.cu restates numpy's legacy global generator (MT19937 + polar Gaussian with a
    cached deviate): same deviates, same state afterwards, for odd/even counts and a pending cache."""
    from cna_b200 import _lib
    for seed, count, pre in [(0, 100001, 0), (5, 7, 3), (1, 250000, 1), (2, 1, 0), (3, 2, 1), (4, 311, 2),
                             (6, 624 * 3, 0), (7, 155, 0), (8, 156, 0), (9, 157, 1)]:
        np.random.seed(seed)
        np.random.randn(pre)
        want, tail_w, u_w = np.random.randn(count), np.random.randn(5), np.random.rand(3)
        np.random.seed(seed)
        np.random.randn(pre)
        got, tail_g, u_g = _lib.host_randn(count, n_threads=3), np.random.randn(5), np.random.rand(3)
        np.testing.assert_array_equal(got, want)
        np.testing.assert_array_equal(tail_g, tail_w)
        np.testing.assert_array_equal(u_g, u_w)
    np.random.seed(11)
    want = np.concatenate([np.random.randn(k) for k in range(1, 120)])
    np.random.seed(11)
    got = np.concatenate([_lib.host_randn(k) for k in range(1, 120)])
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("n,nb,num", [(200, 4, 2000), (50, 1, 333), (333, 7, 257), (12, 12, 40)])
def test_native_permutations_are_bit_exact(n, nb, num):
    """Permutation index matrices from the native engine == the reference's numpy call sequence
    (_stats.py:8-16 and :20-32), including the generator state left behind."""
    from cna_b200.tl import _stats
    rng = np.random.default_rng(n + nb)
    B = rng.integers(0, nb, n)
    B[:nb] = np.arange(nb)
    np.random.seed(n)
    want, after_w = _stats.conditional_permutation_indices(B, num), np.random.randn(4)
    np.random.seed(n)
    got, after_g = _stats.conditional_permutation_matrix(B, num), np.random.randn(4)
    assert got.dtype == np.int32 and got.shape == (num, n)
    np.testing.assert_array_equal(got, want.T)
    np.testing.assert_array_equal(after_g, after_w)
    G = np.arange(n) // 2
    Y = (G % 3 == 0).astype(float)
    np.random.seed(n + 1)
    want = _stats.grouplevel_permutation_indices(G, Y, num)
    np.random.seed(n + 1)
    got = _stats.grouplevel_permutation_matrix(G, Y, num)
    np.testing.assert_array_equal(got, want.T)
    # phenotype that varies within a donor: the reference prints an error and returns None
    assert _stats.grouplevel_permutation_matrix(G, rng.normal(size=n), num) is None


def test_leading_eigenpairs_match_full_svd():
    """svd_of_gram(top=k): same leading singular values and the same projectors U_k U_k^T as the
    reference's full np.linalg.svd of the Gram (_nam.py:105); column signs are free."""
    from cna_b200.tl._nam import svd_of_gram
    rng = np.random.default_rng(0)
    X = rng.normal(size=(120, 5000))
    X -= X.mean(axis=0)
    G = X @ X.T
    U, svs, _ = np.linalg.svd(G)
    Ut, st, _ = svd_of_gram(G, top=16)
    np.testing.assert_allclose(st[:16], svs[:16], rtol=1e-12)
    assert (st[16:] == 0).all() and (Ut[:, 16:] == 0).all()
    for k in (1, 4, 16):
        np.testing.assert_allclose(Ut[:, :k] @ Ut[:, :k].T, U[:, :k] @ U[:, :k].T, atol=1e-10)
    Uf, sf, _ = svd_of_gram(G)
    np.testing.assert_array_equal(sf, svs)


def test_obs_to_sample_matches_reference_semantics():
    """cna.ut.obs_to_sample (utils/multisample.py:4-11): groupby-aggregate, index = ids in order of
    first appearance; against the reference itself when it is mounted."""
    import cna_b200 as cna
    from tests.golden import cases
    d = cases.demo_anndata()
    got = cna.ut.obs_to_sample(d, ["case", "male", "batch"], "id")
    assert list(got.index) == list(d.obs["id"].unique()) and list(got.columns) == ["case", "male", "batch"]
    np.testing.assert_array_equal(got["case"].to_numpy(), d.obs.groupby("id")["case"].mean().reindex(got.index).to_numpy())
    one = cna.ut.obs_to_sample(d, "batch", "id", aggregate="max")
    assert list(one.columns) == ["batch"]
    ref_dir = "/root/reference/src/cna/utils"
    if os.path.isdir(ref_dir):
        import importlib.util
        spec = importlib.util.spec_from_file_location("_ref_multisample", os.path.join(ref_dir, "multisample.py"))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        pd.testing.assert_frame_equal(got, ref.obs_to_sample(d, ["case", "male", "batch"], "id"))


def test_obs_column_adoption_semantics():
    """association() hands its per-cell result buffers to ``data.obs`` without a second copy: the column
    must hold the values, keep the buffer alive, survive the next call's overwrite of the same key, and
    work for any obs index (non-unique labels included)."""
    from cna_b200.tl._association import _adopt_column
    N = 5000
    for index in (pd.RangeIndex(N), pd.Index(np.arange(N) % 7), pd.Index([f"c{i}" for i in range(N)])):
        obs = pd.DataFrame({"id": np.arange(N) % 50}, index=index)

        def fresh(value):
            return (torch.arange(N, dtype=torch.float64) + value).numpy()  # the tensor itself goes out of scope

        _adopt_column(obs, "coef", fresh(0.0))
        assert obs["coef"].dtype == np.float64
        np.testing.assert_array_equal(obs["coef"].to_numpy(), np.arange(N))
        snapshot = obs["coef"]
        _adopt_column(obs, "coef", fresh(7.0))  # the next call overwrites the key with a new buffer
        np.testing.assert_array_equal(obs["coef"].to_numpy(), np.arange(N) + 7.0)
        np.testing.assert_array_equal(snapshot.to_numpy(), np.arange(N))
        nan_col = np.full(N, np.nan)
        _adopt_column(obs, "coef_fdr", nan_col)
        assert obs["coef_fdr"].isna().all() and list(obs.columns) == ["id", "coef", "coef_fdr"]
        obs.iloc[0, obs.columns.get_loc("coef")] = -1.0  # user-side writes keep working
        assert obs["coef"].iloc[0] == -1.0 and obs["coef"].iloc[1] == 8.0


def test_tile_plan_is_consistent():
    """TilePlan (tl/_graph.py) of the shared-memory-staged diffusion step: every edge's position points
    at its own source row in its tile's list, lists are padded to multiples of 4 with valid rows, tiles
    that would exceed the source capacity are halved, a row no tile can hold is refused."""
    import scipy.sparse as sp
    from cna_b200 import _lib
    from cna_b200.tl import _graph
    tile_rows, cap = _lib.diffuse_tile_limits()
    n = 3000
    A = sp.random(n, n, density=0.006, format="lil", random_state=1)
    A[7, :700] = 1.0
    A[8, 300:1100] = 2.0
    A = A.tocsr()
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a))  # noqa: E731
    plan = _graph.TilePlan(t(A.indptr.astype(np.int32)), t(A.indices.astype(np.int32)), t(A.data.astype(np.float32)), n)
    tr, tu, us, ep = plan.tile_row.numpy(), plan.tile_u.numpy(), plan.usrc.numpy(), plan.epair.numpy()
    assert tr[0] == 0 and tr[-1] == n and (np.diff(tr) > 0).all() and np.diff(tr).max() <= tile_rows
    assert (np.diff(tu) % 4 == 0).all() and np.diff(tu).max() <= cap and plan.n_tiles == len(tr) - 1
    assert ((us >= 0) & (us < n)).all()
    assert (ep[:, 1].view(np.float32) == A.data.astype(np.float32)).all()
    for k in range(plan.n_tiles):
        e0, e1 = A.indptr[tr[k]], A.indptr[tr[k + 1]]
        pos = ep[e0:e1, 0] // 128
        assert (ep[e0:e1, 0] % 128 == 0).all() and (pos < tu[k + 1] - tu[k]).all()
        assert (us[tu[k] + pos] == A.indices[e0:e1]).all()
    B = A.tolil()
    B[9, :] = 1.0  # more distinct sources than any tile can stage
    B = B.tocsr()
    with pytest.raises(_lib.CnaError):
        _graph.TilePlan(t(B.indptr.astype(np.int32)), t(B.indices.astype(np.int32)), t(B.data.astype(np.float32)), n)


def test_namespace_mirrors_the_reference_package():
    """``import cna_b200 as cna`` offers the reference's public names (src/cna/__init__.py:1-3,
    tools/__init__.py, plotting/__init__.py, utils/__init__.py); the plotting helpers import their optional
    stack on use and say so when it is missing."""
    import inspect

    import cna_b200 as cna
    for name in ("association", "nam", "svd_nam", "diffuse", "diffuse_stepwise"):
        assert callable(getattr(cna.tl, name))
    assert callable(cna.ut.obs_to_sample)
    for name, params in (("umap_ncorr", ["data", "fdr_thresh", "key"]),
                         ("umap_overlay", ["data", "mask", "key", "scatter0", "scatter1", "ax", "noframe"]),
                         ("violinplot", ["data", "stratification", "key", "ax", "cmap"])):
        fn = getattr(cna.pl, name)
        assert list(inspect.signature(fn).parameters)[:len(params)] == params
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="optional dependency 'matplotlib'"):
            cna.pl.violinplot(None, "cluster")


def test_all_samples_selected_predicate():
    """The cheap eligibility test for taking the QC statistic out of the residualisation pass
    (tl/_association.py:_all_samples_selected): every sample selected, in label order, complete inputs."""
    from cna_b200.tl._association import _all_samples_selected as ok
    labels = pd.Index(np.arange(12))
    y = pd.Series(np.arange(12) % 2, index=labels, dtype=float)
    b = pd.Series(np.arange(12) % 3, index=labels)
    covs = pd.DataFrame({"age": np.linspace(0, 1, 12)}, index=labels)
    assert ok(labels, y, b, covs) and ok(labels, y, b, None)
    assert not ok(labels, y, None, covs)                                  # no batches: no QC at all
    assert not ok(labels, y.iloc[::-1], b.iloc[::-1], None)               # phenotype in another order
    assert not ok(labels, y.iloc[:-1], b.iloc[:-1], None)                 # a sample without phenotype
    y_nan = y.copy()
    y_nan.iloc[3] = np.nan
    assert not ok(labels, y_nan, b, None)                                 # NaN phenotype filters a sample
    c_nan = covs.copy()
    c_nan.iloc[5, 0] = np.nan
    assert not ok(labels, y, b, c_nan)
    assert not ok(labels, y, pd.Series(np.zeros(12), index=labels), None)  # a single batch
    assert not ok(labels, y, b.astype(str), None)                         # non-numeric batches take the general route
    assert not ok(pd.Index(np.arange(13)), y, b, None)                    # data has a sample the phenotype lacks


def test_series_key_tracks_content():
    from cna_b200.tl._nam import _series_key
    a = pd.Series([0, 1, 0, 2], index=[3, 1, 2, 0])
    assert _series_key(a) == _series_key(a.copy())
    assert _series_key(a) != _series_key(pd.Series([0, 1, 0, 3], index=[3, 1, 2, 0]))
    assert _series_key(a) != _series_key(pd.Series([0, 1, 0, 2], index=[3, 1, 0, 2]))
    s = pd.Series(["x", "y"], index=["a", "b"])
    assert _series_key(s) == _series_key(s.copy()) and _series_key(s) != _series_key(s.iloc[::-1])


def test_resid_tables_fit_mirrors_the_kernel_budget():
    """tl/_association.py:_resid_tables_fit restates the shared-memory test of cna_resid_pass's
    linear-functional kernel (csrc/nam_pass.cu: (m1p + r + 1) ldn + nbk r + r + 1 + warps R (m1p + 3) doubles
    + ldn ints <= 100 KiB): the benchmark and golden shapes fit, large designs take the general route."""
    from cna_b200.tl._association import _resid_tables_fit as fits
    assert fits(200, 5, 4) and fits(100, 5, 4) and fits(40, 11, 10) and fits(50, 6, 5)
    assert not fits(256, 40, 16) and not fits(300, 5, 4)
    src = open(os.path.join(ROOT, "cna_b200", "csrc", "nam_pass.cu")).read()
    assert "size_t(m1p + a.r + 1) * ldn + size_t(tb.nbk) * a.r + a.r + 1" in src and "smem <= 100 * 1024" in src
