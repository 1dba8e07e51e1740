"""GPU parity of the tcgen05 kernels (gemm_tc.cu) through the C-ABI: the fp16 hi/lo split, the
tensor-core Gram, X.B with store epilogue and the null GEMM with histogram epilogue, against
float64 numpy on the same fp32 inputs.  Tolerances are fp32-grade (the split keeps 22 bits)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from cna_b200 import _lib
    _lib.load()
    return _lib


def _dev(a, dtype=None):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda", dtype=dtype)


def _padded(a32, mult=8):
    import torch
    rows, cols = a32.shape
    ld = (cols + mult - 1) // mult * mult
    t = torch.zeros((rows, ld), dtype=torch.float32, device="cuda")
    t[:, :cols] = _dev(a32)
    return t


@pytest.mark.parametrize("rows,cols", [(1000, 200), (37, 5), (4, 300)])
def test_split_f16(lib, rows, cols):
    rng = np.random.default_rng(rows)
    a = (rng.normal(size=(rows, cols)) * np.exp(rng.normal(size=(rows, cols)) * 2)).astype(np.float32)
    t = _padded(a)
    p = lib.split_f16(t, cols)
    rec = (p.hi.float() + p.lo.float()).cpu().numpy()
    assert rec.shape == (rows, p.ld)
    np.testing.assert_allclose(rec[:, :cols], a, rtol=2.0 ** -21, atol=2.0 ** -24)
    assert (rec[:, cols:] == 0).all()
    pt = lib.split_f16(t, cols, transpose=True)
    rec = (pt.hi.float() + pt.lo.float()).cpu().numpy()
    assert rec.shape == (cols, pt.ld)
    np.testing.assert_allclose(rec[:, :rows], a.T, rtol=2.0 ** -21, atol=2.0 ** -24)
    assert (rec[:, rows:] == 0).all()


@pytest.mark.parametrize("N,n,n_out", [(1000, 200, 200), (130, 37, 5), (5000, 330, 300), (128, 16, 256),
                                       (20000, 100, 1000), (257, 50, 257)])
def test_right_multiply_tc(lib, N, n, n_out):
    import torch
    rng = np.random.default_rng(N + n)
    x = rng.normal(size=(N, n)).astype(np.float32)
    b = rng.normal(size=(n, n_out)).astype(np.float32)
    xp = lib.split_f16(_padded(x), n)
    btp = lib.split_f16(_padded(b), n_out, transpose=True)  # planes of B^T: [n_out x n]
    ld_out = (n_out + 3) // 4 * 4
    out = torch.full((N, ld_out), 7.0, dtype=torch.float32, device="cuda")
    lib.right_multiply_tc(xp, n, btp, n_out, out)
    got = out.cpu().numpy()
    want = x.astype(np.float64) @ b.astype(np.float64)
    scale = np.sqrt((x.astype(np.float64) ** 2).sum(1))[:, None] * np.sqrt((b.astype(np.float64) ** 2).sum(0))[None, :]
    err = np.abs(got[:, :n_out] - want) / scale
    assert err.max() < 2e-6, err.max()
    assert (got[:, n_out:] == 7.0).all()  # padding columns untouched
    # agrees with the CUDA-core kernel to fp32 rounding
    ref = torch.empty((N, ld_out), dtype=torch.float32, device="cuda")
    lib.right_multiply(_padded(x), n, _padded(b, 4), n_out, ref)
    assert (np.abs(ref.cpu().numpy()[:, :n_out] - want) / scale).max() < 2e-6


@pytest.mark.parametrize("N,n", [(5000, 200), (700, 50), (100000, 200), (3000, 129), (513, 16), (40000, 256),
                                 (6000, 257), (20000, 330), (9000, 400), (50000, 500), (2000, 512), (100, 300)])
def test_gram_tc(lib, N, n):
    import torch
    rng = np.random.default_rng(N + n)
    x = rng.normal(size=(N, n)).astype(np.float32)
    x[:, 0] += 1.0  # a column with non-zero mean: all-positive cross terms
    x[7] = 0.0
    xt = _padded(x)
    xp = lib.split_f16(xt, n)
    G = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    G[0, 0] = 5.0  # the kernel accumulates into G
    lib.gram_tc(xp, n, G)
    x64 = x.astype(np.float64)
    want = x64.T @ x64
    want[0, 0] += 5.0
    got = G.cpu().numpy()
    d = np.sqrt(np.diag(want))
    err = np.abs(got - want) / (d[:, None] * d[None, :])
    assert err.max() < 3e-6, err.max()
    # diagonal (all-positive sums): systematic truncation bias must stay well inside 1e-5
    rel = np.abs(np.diag(got) - np.diag(want)) / np.diag(want)
    assert rel.max() < 3e-6, rel.max()
    # deterministic (no atomics)
    G2 = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    G2[0, 0] = 5.0
    lib.gram_tc(xp, n, G2)
    assert torch.equal(G, G2)


@pytest.mark.parametrize("N,n,Kl", [(4100, 200, 100), (1000, 100, 1000), (30000, 200, 333), (2000, 330, 64)])
def test_null_hist_tc(lib, N, n, Kl):
    import torch
    from cna_b200.tl import _stats
    rng = np.random.default_rng(N + Kl)
    x = rng.normal(size=(N, n)).astype(np.float32)
    x[3] = 0.0
    yc = rng.normal(size=(n, Kl)).astype(np.float32)
    x64, y64 = x.astype(np.float64), yc.astype(np.float64)
    z2 = (x64 @ y64 / n) ** 2
    mx = np.sqrt(z2.max()) * 0.8
    thr = np.arange(mx / 4, mx, mx / 400)
    edges = _stats.threshold_edges(thr)
    xp = lib.split_f16(_padded(x), n)
    ytp = lib.split_f16(_padded(yc, 4), Kl, transpose=True)
    hist = torch.zeros(len(thr), dtype=torch.int64, device="cuda")
    hist[0] = 3  # the kernel accumulates
    lib.null_hist_tc(xp, n, ytp, Kl, _dev(edges), float(edges[0]), hist)
    hs = hist.cpu().numpy()
    hs[0] -= 3
    tails_sum = _stats.tails_from_hist(hs)  # summed over the Kl nulls
    lo = np.stack([(z2 >= e * (1 + 1e-5)).sum(0) for e in edges], axis=1)
    hi = np.stack([(z2 >= e * (1 - 1e-5)).sum(0) for e in edges], axis=1)
    assert (tails_sum >= lo.sum(0)).all() and (tails_sum <= hi.sum(0)).all()
    assert tails_sum.sum() > 0
    # same counts as the CUDA-core kernel up to the same knife edge
    ycp = torch.zeros((xp.ld if False else (n + 7) // 8 * 8, (Kl + 3) // 4 * 4), dtype=torch.float32, device="cuda")
    ycp[:n, :Kl] = _dev(yc)
    hist2 = torch.zeros((Kl, len(thr)), dtype=torch.int32, device="cuda")
    lib.null_hist(_padded(x), n, ycp, Kl, _dev(edges), float(edges[0]), hist2)
    t2 = _stats.tails_from_hist(hist2.cpu().numpy().astype(np.int64))
    assert (t2 >= lo).all() and (t2 <= hi).all()
    # the summed table gives the same FDR as the per-null one
    h2 = hist2.cpu().numpy().astype(np.int64)
    rank_hist = h2[0] + 1
    np.testing.assert_allclose(_stats.fdr_from_counts(h2.sum(0), rank_hist, n_null=Kl),
                               _stats.fdr_from_counts(h2, rank_hist), rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("n,k,kind", [(200, 16, "gram"), (50, 4, "gram"), (500, 40, "gram"), (216, 17, "gram"),
                                      (230, 10, "gram"), (129, 129 - 1, "gram"), (120, 12, "clustered"),
                                      (37, 5, "lowrank"), (3, 2, "gram"), (2, 1, "gram"), (300, 150, "gram")])
def test_sym_eig_top_vs_lapack(lib, n, k, kind):
    """cna_sym_eig_top (device Householder + multisection + inverse iteration) against numpy's eigh:
    eigenvalues to 1e-12 of the largest, projectors onto every leading subspace that is separated from
    the rest to 1e-9, unit orthogonal vectors, small residual.  Replaces _nam.py:105 for the columns
    _association.py:35-42 reads."""
    import torch
    rng = np.random.default_rng(n * 1000 + k)
    if kind == "gram":
        X = rng.normal(size=(20 * n, n)) * np.linspace(3.0, 1.0, n)
        G = X.T @ X
    elif kind == "lowrank":  # rank 10 < n: trailing eigenvalues are zero up to rounding
        X = rng.normal(size=(10, n))
        G = X.T @ X
    else:  # near-multiple leading eigenvalues: the re-orthogonalisation inside clusters
        Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
        ev = np.linspace(1.0, 2.0, n)
        ev[-3:] = [5.0, 5.0 + 1e-9, 5.0 + 2e-9]
        ev[-6:-3] = [4.0, 4.0 + 1e-13, 4.0 + 3e-13]
        G = (Q * ev) @ Q.T
    G = G + 1e-9 * rng.normal(size=(n, n)) * np.abs(G).max() * 1e-7  # not exactly symmetric, like the device Gram
    Gs = (G + G.T) / 2
    G_d = torch.as_tensor(G, device="cuda")
    w_d = torch.empty(k, dtype=torch.float64, device="cuda")
    ut_d = torch.empty((k, n), dtype=torch.float64, device="cuda")
    de_d = torch.empty(2 * n + 4, dtype=torch.float64, device="cuda")
    lib.sym_eig_top(G_d, k, w_d, ut_d, de_d)
    torch.cuda.synchronize()
    w, Ut, de = w_d.cpu().numpy(), ut_d.cpu().numpy(), de_d.cpu().numpy()
    w0, U0 = np.linalg.eigh(Gs)
    w0, U0 = w0[::-1], U0[:, ::-1]
    scale = abs(w0[0])
    # the tridiagonal form has the same spectrum
    import scipy.linalg as sl
    if n > 1:
        wt = sl.eigvalsh_tridiagonal(de[:n], de[n:2 * n - 1])[::-1]
        np.testing.assert_allclose(wt, w0, rtol=0, atol=1e-12 * scale)
    np.testing.assert_allclose(w, w0[:k], rtol=0, atol=1e-12 * scale)
    np.testing.assert_allclose(Ut @ Ut.T, np.eye(k), rtol=0, atol=1e-10)
    resid = Gs @ Ut.T - Ut.T * w
    assert np.abs(resid).max() <= 1e-10 * scale
    for j in range(1, k + 1):  # every leading subspace with a gap behind it
        gap = (w0[j - 1] - w0[j]) / scale if j < n else 1.0
        if gap > 1e-6:
            P, P0 = Ut[:j].T @ Ut[:j], U0[:, :j] @ U0[:, :j].T
            assert np.abs(P - P0).max() <= 1e-9 / gap * 1e-3 + 1e-11, (j, gap)
