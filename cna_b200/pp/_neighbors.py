"""kNN graph construction on the GPU: what ``scanpy.pp.neighbors(adata, n_neighbors=k)`` leaves in
``.obsp['connectivities']`` / ``.obsp['distances']`` / ``.uns['neighbors']`` (method 'umap', metric
'euclidean'), which is what the reference reads at ``src/cna/tools/_nam.py:12-19``.

Differences from scanpy, all deliberate: the search is EXACT (a tiled brute-force kernel,
``cna_knn_bruteforce_range``: candidates stream through shared memory, nothing of size N x N is ever
formed) where scanpy switches to approximate NN-descent above 4096 cells; distances are computed in
float32 like pynndescent's; no PCA is run here (pass the representation, ``use_rep``).  The weight
construction follows umap-learn's ``fuzzy_simplicial_set`` (smooth-kNN distances by bisection,
``local_connectivity=1``, symmetrisation by the probabilistic t-conorm ``P + P^T - P o P^T``) and is
pinned against the reference's own fixture: from the kNN distances stored in ``demo/data.h5ad`` it
reproduces the stored connectivities to 2e-6 (``tests/test_host_logic.py``).
"""
import math

import numpy as np
import scipy.sparse as sp
import torch


def knn(points, k, queries=None, block=1 << 20):
    """Exact k nearest neighbours (self excluded) of ``points`` [N x d] (d <= 64) on the current CUDA
    device.  Returns (idx int64 [nq x k], dist float64 [nq x k]) sorted by distance, for the query rows
    ``queries`` = (q0, q1) (default: all).  The search runs in blocks of ``block`` queries."""
    from .. import _lib
    pts = torch.as_tensor(np.asarray(points) if not torch.is_tensor(points) else points,
                          dtype=torch.float32, device="cuda").contiguous()
    n = pts.shape[0]
    q0, q1 = (0, n) if queries is None else queries
    idx_parts, d_parts = [], []
    for b0 in range(q0, q1, block):
        i, d2 = _lib.knn_bruteforce(pts, k, queries=(b0, min(b0 + block, q1)))
        idx_parts.append(i)
        d_parts.append(d2)
    idx = torch.cat(idx_parts) if len(idx_parts) != 1 else idx_parts[0]
    d2 = torch.cat(d_parts) if len(d_parts) != 1 else d_parts[0]
    return idx, d2.double().sqrt_()


def fuzzy_simplicial_set(idx, dist, n_iter=64, rows=None, n_total=None):
    """UMAP's smooth-kNN-distance weights, symmetrised by probabilistic t-conorm.

    idx/dist: [N, k-1] neighbour indices / distances (self excluded), torch tensors (any device) or
    numpy arrays.  Returns scipy CSR float64 with sorted indices; with ``rows`` = (r0, r1) only that
    block of rows, as an (r1 - r0) x N matrix (what one rank of a sharded run ingests)."""
    idx = torch.as_tensor(idx)
    dist = torch.as_tensor(dist, dtype=torch.float64, device=idx.device)
    idx = idx.long()
    N, km1 = idx.shape
    if n_total is not None and n_total != N:
        raise ValueError("fuzzy_simplicial_set needs the neighbour lists of all N cells (the symmetrisation "
                         "of a row block reads the lists of its neighbours)")
    target = math.log2(km1 + 1)
    rho = dist[:, :1]
    gap = (dist - rho).clamp_(min=0)
    lo = torch.zeros(N, 1, dtype=torch.float64, device=idx.device)
    hi = torch.full((N, 1), float("inf"), dtype=torch.float64, device=idx.device)
    mid = torch.ones(N, 1, dtype=torch.float64, device=idx.device)
    for _ in range(n_iter):
        psum = torch.exp(-gap / mid).sum(dim=1, keepdim=True)
        too_big = psum > target
        hi = torch.where(too_big, mid, hi)
        lo = torch.where(too_big, lo, mid)
        mid = torch.where(torch.isinf(hi), mid * 2, (lo + hi) / 2)
    p = torch.exp(-gap / mid).reshape(-1)
    i = torch.arange(N, device=idx.device).repeat_interleave(km1)
    j = idx.reshape(-1)
    key = torch.cat([i * N + j, j * N + i])
    val = torch.cat([p, p])
    key, order = torch.sort(key)
    val = val[order]
    ukey, inv = torch.unique_consecutive(key, return_inverse=True)
    s1 = torch.zeros(len(ukey), dtype=torch.float64, device=idx.device).index_add_(0, inv, val)
    s2 = torch.zeros(len(ukey), dtype=torch.float64, device=idx.device).index_add_(0, inv, val * val)
    w = s1 - (s1 * s1 - s2) / 2  # p + q - p.q for mutual pairs, p otherwise
    r0, r1 = (0, N) if rows is None else rows
    if rows is not None:  # keys are sorted by row: the block is one slice
        e0, e1 = torch.searchsorted(ukey, torch.tensor([r0 * N, r1 * N], device=ukey.device)).tolist()
        ukey, w = ukey[e0:e1], w[e0:e1]
    row = (ukey // N - r0).cpu().numpy()
    cols = (ukey % N).cpu().numpy().astype(np.int32)
    indptr = np.zeros(r1 - r0 + 1, dtype=np.int64)
    np.cumsum(np.bincount(row, minlength=r1 - r0), out=indptr[1:])
    return sp.csr_matrix((w.cpu().numpy(), cols, indptr.astype(np.int32)), shape=(r1 - r0, N))


def _representation(data, use_rep):
    if use_rep is not None and not isinstance(use_rep, str):
        return np.asarray(use_rep)
    obsm = getattr(data, "obsm", None)
    if use_rep is None:
        if obsm is not None and "X_pca" in obsm:
            return np.asarray(obsm["X_pca"])
        use_rep = "X"
    if use_rep == "X":
        X = getattr(data, "X", None)
        if X is None:
            raise ValueError("neighbors: data has neither .obsm['X_pca'] nor .X; pass use_rep=<array>")
        return np.asarray(X.todense() if sp.issparse(X) else X)
    if obsm is None or use_rep not in obsm:
        raise ValueError(f"neighbors: .obsm has no '{use_rep}'")
    return np.asarray(obsm[use_rep])


def neighbors(data, n_neighbors=15, use_rep=None, key_added=None, rows=None, copy=False):
    """``scanpy.pp.neighbors(data, n_neighbors=n_neighbors, use_rep=use_rep, key_added=key_added)`` with an
    exact GPU search.  Writes ``data.obsp['connectivities']``, ``data.obsp['distances']`` and
    ``data.uns['neighbors']`` (or ``'<key_added>_connectivities'`` ... like scanpy) and returns None, or
    returns (connectivities, distances) when ``copy``.

    ``use_rep``: a key of ``.obsm``, ``'X'``, or an array [N x d], d <= 64 (default: ``.obsm['X_pca']`` if
    present, else ``.X``).  ``rows`` = (r0, r1): keep only that block of rows of both matrices (the kNN search
    still covers all cells: the symmetrisation needs every neighbour list) — the form one rank of
    ``cna_b200.sharded`` ingests; sets ``data.row_block``."""
    X = _representation(data, use_rep)
    if X.ndim != 2 or X.shape[1] > 64:
        raise ValueError(f"neighbors: the representation must be [cells x d] with d <= 64 (got {X.shape}); "
                         "reduce it first (scanpy.pp.pca)")
    N = X.shape[0]
    if not 2 <= n_neighbors <= 65 or n_neighbors > N:
        raise ValueError("neighbors: n_neighbors must be in [2, 65] and at most the number of cells")
    idx, dist = knn(X, n_neighbors - 1)
    conn = fuzzy_simplicial_set(idx, dist, rows=rows)
    r0, r1 = (0, N) if rows is None else rows
    ii = idx[r0:r1].cpu().numpy().astype(np.int32)
    dd = dist[r0:r1].cpu().numpy()
    order = np.argsort(ii, axis=1, kind="stable")  # sorted column indices inside a row, like scipy's CSR
    ii, dd = np.take_along_axis(ii, order, 1), np.take_along_axis(dd, order, 1)
    indptr = np.arange(0, (r1 - r0) * (n_neighbors - 1) + 1, n_neighbors - 1, dtype=np.int32)
    dmat = sp.csr_matrix((dd.reshape(-1), ii.reshape(-1), indptr), shape=(r1 - r0, N))
    if copy:
        return conn, dmat
    ckey = "connectivities" if key_added is None else f"{key_added}_connectivities"
    dkey = "distances" if key_added is None else f"{key_added}_distances"
    if getattr(data, "obsp", None) is None:
        data.obsp = {}
    data.obsp[ckey], data.obsp[dkey] = conn, dmat
    if getattr(data, "uns", None) is None:
        data.uns = {}
    data.uns["neighbors" if key_added is None else key_added] = {
        "connectivities_key": ckey, "distances_key": dkey,
        "params": {"n_neighbors": int(n_neighbors), "method": "umap", "metric": "euclidean", "random_state": 0,
                   "search": "exact (cna_b200)"}}
    if rows is not None:
        data.row_block = (r0, r1)
    return None
