"""Synthetic multi-sample single-cell datasets shaped like the reference's demo.

There is no scanpy/umap/anndata in this image and no network for real datasets, so the parity tests
and ``bench.py`` build their own AnnData-like inputs: a latent Gaussian mixture whose cluster
proportions differ between cases and controls, an exact kNN graph and UMAP's fuzzy-simplicial-set
weights symmetrised as ``A = P + P^T - P o P^T`` — i.e. what ``scanpy.pp.neighbors`` would leave in
``.obsp['connectivities']`` (float64 CSR, sorted indices, zero diagonal, weights in (0, 1]).

The kNN search uses scikit-learn on the CPU for test-sized inputs and the library's brute-force
CUDA kernel (``cna_knn_bruteforce``) for benchmark-sized inputs when a GPU is present; the weight
construction runs in torch on whichever device is available.  None of this is on the timed path.
"""
import math

import numpy as np
import pandas as pd
import scipy.sparse as sp
import torch


class AnnDataLike:
    """Duck-typed AnnData: the hot path only touches ``.obs`` and ``.obsp['connectivities']``
    (reference ``_nam.py:19,51``; ``_association.py:228-237``)."""

    def __init__(self, obs, connectivities):
        self.obs = obs
        self.obsp = {"connectivities": connectivities}

    @property
    def n_obs(self):
        return len(self.obs)

    def __repr__(self):
        a = self.obsp["connectivities"]
        return f"AnnDataLike(n_obs={self.n_obs}, nnz={a.nnz}, obs={list(self.obs.columns)})"


def _knn_cpu(points, k):
    from sklearn.neighbors import NearestNeighbors
    nn = NearestNeighbors(n_neighbors=k, algorithm="auto").fit(points)
    dist, idx = nn.kneighbors(points)
    # drop self (first hit; duplicates are vanishingly unlikely with continuous data)
    return idx[:, 1:].astype(np.int64), dist[:, 1:].astype(np.float64)


def _knn_gpu(points, k, comm=None):
    """Exact kNN on the GPU; with a ``comm`` (cna_b200.sharded.Comm) every rank searches its own block
    of queries and the blocks are all-gathered (brute force is O(N^2): 10M points take ~6 min on one
    B200, ~45 s split over eight)."""
    from . import _lib
    pts = torch.as_tensor(points, dtype=torch.float32, device="cuda").contiguous()
    if comm is None or comm.world == 1:
        idx, d2 = _lib.knn_bruteforce(pts, k - 1)
        return idx, d2.double().sqrt_()
    n = pts.shape[0]
    per = (n + comm.world - 1) // comm.world
    q0, q1 = min(comm.rank * per, n), min((comm.rank + 1) * per, n)
    idx, d2 = _lib.knn_bruteforce(pts, k - 1, queries=(q0, q1))
    pad_i = torch.zeros((per, k - 1), dtype=torch.int32, device=pts.device)
    pad_d = torch.zeros((per, k - 1), dtype=torch.float32, device=pts.device)
    pad_i[: q1 - q0], pad_d[: q1 - q0] = idx.to(torch.int32), d2
    idx = comm.all_gather_rows(pad_i)[:n].long()
    d2 = comm.all_gather_rows(pad_d)[:n]
    return idx, d2.double().sqrt_()


from .pp._neighbors import fuzzy_simplicial_set  # noqa: E402,F401  (the product implementation)


def make_dataset(n_cells=10000, n_samples=50, k=15, dim=None, seed=0, n_clusters=12, n_batches=4,
                 ragged=False, knn="auto", device=None, comm=None, row_block=False):
    """Returns (AnnDataLike, sample_meta DataFrame with columns case / batch / age).

    Cells are stored sample-contiguous like the reference's demo; ``ragged`` draws unequal cells
    per sample.  ``obs`` has the column ``id`` (sample id per cell).  With a ``comm`` the kNN search is
    split over the ranks (every rank ends with the same data); ``row_block=True`` then keeps only this
    rank's block of rows of the graph on the host (``.obsp['connectivities']`` is (rows of the block) x
    N, ``.row_block`` = (r0, r1)): what ``cna_b200.sharded.shard_to_device`` needs and nothing more."""
    rng = np.random.default_rng(seed)
    if dim is None:
        dim = 20 if n_cells <= 200_000 else 6
    if ragged:
        w = rng.uniform(0.5, 1.5, n_samples)
        counts = np.maximum((w / w.sum() * n_cells).astype(np.int64), 2)
        counts[-1] += n_cells - counts.sum()
    else:
        counts = np.full(n_samples, n_cells // n_samples, dtype=np.int64)
        counts[: n_cells - counts.sum()] += 1
    sid = np.repeat(np.arange(n_samples), counts)
    case = (np.arange(n_samples) >= n_samples / 2).astype(np.float64)
    centres = rng.normal(0, 4.0, (n_clusters, dim))
    logits = rng.normal(0, 0.3, (n_samples, n_clusters))
    logits[:, 0] += 0.8 * case
    probs = np.exp(logits)
    probs /= probs.sum(axis=1, keepdims=True)
    cum = np.cumsum(probs, axis=1)
    u = rng.random(n_cells)
    cluster = (u[:, None] > cum[sid]).sum(axis=1).clip(max=n_clusters - 1)
    points = centres[cluster] + rng.normal(0, 1.0, (n_cells, dim))

    use_gpu = torch.cuda.is_available() if knn == "auto" else knn == "gpu"
    if use_gpu and n_cells >= 50_000:
        idx, dist = _knn_gpu(points, k, comm=comm)
    else:
        idx, dist = _knn_cpu(points, k)
        if device is None and torch.cuda.is_available():
            device = "cuda"
        idx = torch.as_tensor(idx, device=device or "cpu")
        dist = torch.as_tensor(dist, device=device or "cpu")
    block = None
    if row_block and comm is not None and comm.world > 1:
        per = (n_cells + comm.world - 1) // comm.world
        block = (min(comm.rank * per, n_cells), min((comm.rank + 1) * per, n_cells))
    A = fuzzy_simplicial_set(idx, dist, rows=block)
    del idx, dist

    index = (pd.Index([f"c{i}" for i in range(n_cells)]) if n_cells <= 200_000
             else pd.RangeIndex(n_cells))
    obs = pd.DataFrame({"id": sid}, index=index)
    meta = pd.DataFrame({
        "case": case,
        "batch": np.tile(np.arange(n_batches), n_samples // n_batches + 1)[:n_samples],
        "age": np.random.default_rng(seed + 1).normal(0, 1, n_samples),
    }, index=pd.Index(np.arange(n_samples), name="id"))
    data = AnnDataLike(obs, A)
    if block is not None:
        data.row_block = block
    return data, meta
