"""The two steps either side of the hot path: ``.h5ad`` ingestion and kNN-graph
construction.  CPU tests cover the reader and the fuzzy-simplicial-set weights against the reference's own
fixture; the GPU tests cover the exact kNN search and the whole ``pp.neighbors`` -> ``tl.association`` chain."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
DEMO = "/root/reference/demo/data.h5ad"


def _demo_knn():
    z = np.load(os.path.join(GOLDEN, "demo_knn.npz"))
    n = len(z["indptr"]) - 1
    km1 = int(z["n_neighbors"]) - 1
    idx = z["indices"].reshape(n, km1).astype(np.int64)
    dist = z["data"].astype(np.float64).reshape(n, km1)
    order = np.argsort(dist, axis=1, kind="stable")  # scanpy stores a row sorted by column, UMAP wants distance order
    return np.take_along_axis(idx, order, 1), np.take_along_axis(dist, order, 1)


def test_fuzzy_simplicial_set_reproduces_the_reference_fixture():
    """From the kNN distances scanpy stored in demo/data.h5ad, cna_b200.pp.fuzzy_simplicial_set rebuilds the
    connectivities stored in the same file (the graph every golden case diffuses over): same sparsity pattern,
    weights to 5e-6 (umap-learn keeps float32 weights and stops its bisection at 1e-5)."""
    from cna_b200.pp import fuzzy_simplicial_set
    idx, dist = _demo_knn()
    g = np.load(os.path.join(GOLDEN, "demo_graph.npz"))
    want = sp.csr_matrix((g["data"], g["indices"], g["indptr"]), shape=(len(idx), len(idx)))
    got = fuzzy_simplicial_set(idx, dist)
    assert got.nnz == want.nnz
    np.testing.assert_array_equal(got.indptr, want.indptr)
    np.testing.assert_array_equal(got.indices, want.indices)
    np.testing.assert_allclose(got.data, want.data, rtol=0, atol=5e-6)
    # one rank's block of rows = the same rows of the whole matrix
    blk = fuzzy_simplicial_set(idx, dist, rows=(2500, 6000))
    assert blk.shape == (3500, len(idx)) and abs(blk - got[2500:6000]).max() == 0


@pytest.mark.skipif(not os.path.exists(DEMO), reason="the reference checkout is not mounted")
def test_read_h5ad_whole_and_row_block():
    """cna_b200.read_h5ad on the reference's bundled file: the legacy uns['neighbors'] graph location
    (_nam.py:17-19), numeric obs columns, and a row block that equals the same rows of the whole matrix."""
    import cna_b200
    d = cna_b200.read_h5ad(DEMO, pin=False)
    g = np.load(os.path.join(GOLDEN, "demo_graph.npz"))
    A = d.obsp["connectivities"]
    assert A.shape == (10000, 10000) and A.dtype == np.float64 and A.indices.dtype == np.int32
    for part in ("data", "indices", "indptr"):
        np.testing.assert_array_equal(getattr(A, part), g[part])
    for col in ("id", "case", "male", "batch"):
        np.testing.assert_array_equal(d.obs[col].to_numpy(), g["obs_" + col])
    b = cna_b200.read_h5ad(DEMO, rows=(1234, 7000), obs_columns=["id"], pin=False)
    B = b.obsp["connectivities"]
    assert b.row_block == (1234, 7000) and B.shape == (7000 - 1234, 10000) and list(b.obs.columns) == ["id"]
    assert abs(B - A[1234:7000]).max() == 0 and B.indptr[0] == 0
    with pytest.raises(KeyError):
        cna_b200.read_h5ad(DEMO, obs_columns=["no_such_column"], pin=False)
    with pytest.raises(ValueError):
        cna_b200.read_connectivities(DEMO, rows=(5, 20000), pin=False)


@pytest.mark.gpu
@pytest.mark.parametrize("n,dim,k", [(3000, 20, 15), (2000, 50, 30), (1500, 6, 10), (700, 64, 5)])
def test_exact_knn_matches_sklearn(n, dim, k):
    """cna_b200.pp.knn (tiled brute force, float32) against scikit-learn's exact search in float64: the same
    neighbour sets wherever the k-th and (k+1)-th distances differ by more than float32 rounding."""
    from sklearn.neighbors import NearestNeighbors

    from cna_b200 import pp
    rng = np.random.default_rng(n + dim)
    X = rng.normal(size=(n, dim)).astype(np.float32)
    idx, dist = pp.knn(X, k - 1)
    idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
    dd, ii = NearestNeighbors(n_neighbors=k + 1).fit(X.astype(np.float64)).kneighbors(X.astype(np.float64))
    np.testing.assert_allclose(dist, dd[:, 1:k], rtol=2e-5, atol=1e-6)
    # rows whose k-th neighbour is not a near tie at float32 resolution (distances concentrate in high dimension)
    clear = (dd[:, k] - dd[:, k - 1]) > 1e-3 * dd[:, k]
    same = np.array([set(a) == set(b) for a, b in zip(idx, ii[:, 1:k])])
    assert same[clear].all() and clear.mean() > 0.3
    overlap = np.mean([len(set(a) & set(b)) / (k - 1) for a, b in zip(idx, ii[:, 1:k])])
    assert overlap > 0.995
    part, _ = pp.knn(X, k - 1, queries=(100, 600))
    np.testing.assert_array_equal(part.cpu().numpy(), idx[100:600])


@pytest.mark.gpu
def test_neighbors_writes_the_scanpy_fields_and_feeds_association():
    """pp.neighbors end to end with the exact search (the weight construction itself is pinned by the CPU test
    above): obsp / uns fields as scanpy writes them, symmetric weights in (0, 1], a row block equal to the
    same rows of the whole graph, and tl.association accepts the result."""
    import pandas as pd

    import cna_b200 as cna
    from cna_b200 import synth
    rng = np.random.default_rng(0)
    n, S = 6000, 30
    sid = np.repeat(np.arange(S), n // S)
    X = rng.normal(size=(n, 20)).astype(np.float32) + (sid[:, None] % 3)
    d = synth.AnnDataLike(pd.DataFrame({"id": sid}), None)
    d.obsp = {}
    d.obsm = {"X_pca": X}
    cna.pp.neighbors(d, n_neighbors=15)
    A, D = d.obsp["connectivities"], d.obsp["distances"]
    assert A.shape == (n, n) and D.shape == (n, n) and (np.diff(D.indptr) == 14).all()
    assert d.uns["neighbors"]["params"]["n_neighbors"] == 15
    assert d.uns["neighbors"]["connectivities_key"] == "connectivities"
    assert abs(A - A.T).max() < 1e-12 and A.data.min() > 0 and A.data.max() <= 1 + 1e-12 and A.diagonal().sum() == 0
    assert (np.diff(A.indptr) >= 14).all()
    conn_blk, dist_blk = cna.pp.neighbors(d, n_neighbors=15, rows=(1000, 2500), copy=True)
    assert abs(conn_blk - A[1000:2500]).max() == 0 and abs(dist_blk - D[1000:2500]).max() == 0
    y = pd.Series((np.arange(S) % 2).astype(float), index=np.arange(S))
    p = cna.tl.association(d, y, "id", Nnull=100, seed=0, nsteps=3)
    assert 0 < p <= 1 and "coef" in d.obs


def test_fuzzy_simplicial_set_properties_on_random_neighbour_lists():
    """Symmetric, zero diagonal, weights in (0, 1], every stored kNN edge present, and the nearest neighbour of
    every cell carries weight 1 (local_connectivity = 1: rho is the distance to it)."""
    from sklearn.neighbors import NearestNeighbors

    from cna_b200.pp import fuzzy_simplicial_set
    rng = np.random.default_rng(2)
    X = rng.normal(size=(1500, 8))
    dist, idx = NearestNeighbors(n_neighbors=11).fit(X).kneighbors(X)
    A = fuzzy_simplicial_set(idx[:, 1:], dist[:, 1:])
    assert A.shape == (1500, 1500) and abs(A - A.T).max() < 1e-15 and A.diagonal().sum() == 0
    assert 0 < A.data.min() and A.data.max() <= 1.0 + 1e-12  # p + q - p q of two ones, up to rounding
    assert (np.asarray(A[np.repeat(np.arange(1500), 10), idx[:, 1:].ravel()]) > 0).all()
    np.testing.assert_allclose(np.asarray(A[np.arange(1500), idx[:, 1]]).ravel(), 1.0, rtol=0, atol=1e-12)
    with pytest.raises(ValueError):
        fuzzy_simplicial_set(idx[:100, 1:], dist[:100, 1:], n_total=1500)


def test_neighbors_argument_checks():
    """pp.neighbors validates before touching the GPU: representation lookup and shape, n_neighbors range."""
    from cna_b200 import pp, synth
    import pandas as pd
    d = synth.AnnDataLike(pd.DataFrame({"id": np.zeros(10, dtype=int)}), None)
    with pytest.raises(ValueError, match="neither"):
        pp.neighbors(d)
    d.obsm = {"X_pca": np.zeros((10, 70))}
    with pytest.raises(ValueError, match="d <= 64"):
        pp.neighbors(d)
    with pytest.raises(ValueError, match="no 'X_other'"):
        pp.neighbors(d, use_rep="X_other")
    d.obsm["X_pca"] = np.zeros((10, 5))
    with pytest.raises(ValueError, match="n_neighbors"):
        pp.neighbors(d, n_neighbors=11)
    with pytest.raises(ValueError, match="n_neighbors"):
        pp.neighbors(d, n_neighbors=1)
