"""Permutation generators and empirical-FDR bookkeeping — host side.

Mirrors ``src/cna/tools/_stats.py`` of the reference.  Permutation *indices* follow exactly the
reference's sequence of legacy numpy RNG calls (``np.random.randn`` + ``np.argsort(axis=0)``): the legacy
MT19937 / polar-Gaussian stream is restated natively, on the device (``cna_perm_draw_device``, the default
engine of ``association``) and on the host (``cna_host_perm_blocks``), and leaves numpy's global generator
in the state the reference's calls would have left it in.  The functions ``*_indices`` keep the literal
numpy call sequence as the in-repo cross-check.  The gather ``Y[bix]`` and everything after it happen on
the device (``cna_perm_stats``).
"""
import numpy as np


def conditional_permutation_indices(B, num):
    """``_stats.py:4-16``: permutation index matrix (n x num) of samples within batches.  Consumes the
    global numpy RNG exactly like the reference: one ``randn(len(batch), num)`` block per batch, in
    sorted batch order."""
    B = np.asarray(B)
    batchind = [np.where(B == b)[0] for b in np.unique(B)]
    ix = np.concatenate([bi[np.argsort(np.random.randn(len(bi), num), axis=0)] for bi in batchind])
    bix = np.zeros((len(B), num), dtype=np.int64)
    bix[np.concatenate(batchind)] = ix
    return bix


def conditional_permutation(B, Y, num):
    """``_stats.py:4-18``."""
    return np.asarray(Y)[conditional_permutation_indices(B, num)]


def grouplevel_permutation_indices(G, Y, num):
    """``_stats.py:20-32`` as sample indices: entry (m, k) is a sample whose phenotype equals the
    value donor-permutation k assigns to sample m.  Returns None (after printing the reference's
    message) when Y is not constant within a donor."""
    G = np.asarray(G)
    Y = np.asarray(Y)
    Gu = np.unique(G)
    rep = np.array([np.where(G == g)[0][0] for g in Gu])  # first sample of each donor
    Yg = Y[rep]
    Gind = np.searchsorted(Gu, G)
    if (Yg[Gind] != Y).any():
        print("ERROR: the value of Y is not identical within each group of samples")
        return None
    ix = np.argsort(np.random.randn(len(Yg), num), axis=0)
    return rep[ix][Gind]


def grouplevel_permutation(G, Y, num):
    ix = grouplevel_permutation_indices(G, Y, num)
    return None if ix is None else np.asarray(Y)[ix]


def _batch_blocks(B):
    B = np.asarray(B)
    batchind = [np.where(B == b)[0] for b in np.unique(B)]
    off = np.zeros(len(batchind) + 1, dtype=np.int32)
    np.cumsum([len(bi) for bi in batchind], out=off[1:])
    return off, np.concatenate(batchind)


class RedoWithHostDraw(Exception):
    """The device draw could not certify an argsort (two keys within rounding of a tie) or ran out of
    stream: the call is repeated with the host engine.  numpy's generator is back in its prior state."""


_DEVICE_DRAW = None


def _device_draw_enabled():
    global _DEVICE_DRAW
    if _DEVICE_DRAW is None:
        import os

        import torch
        _DEVICE_DRAW = bool(torch.cuda.is_available()) and not os.environ.get("CNA_B200_HOST_DRAW")
    return _DEVICE_DRAW


class PermutationDraw:
    """Asynchronous ``conditional_permutation_matrix`` / ``grouplevel_permutation_matrix``.  Default engine:
    the device draw (``cna_perm_draw_device``), queued on a side stream beside the NAM kernels — no host
    threads, and on a cell-axis shard every rank draws the same matrix itself.  ``engine="host"`` (or
    ``CNA_B200_HOST_DRAW=1``): the native host engine on a library-owned thread.  numpy's global generator
    must not be used until ``result()`` / ``cancel()`` has returned (it is advanced there exactly as the
    reference's calls would have advanced it)."""

    def __init__(self, y_std, batches, donorids, num, device=None, engine=None):
        from .. import _lib
        self._map = None
        self._fail = False
        self._dev = None
        use_device = device is not None and (engine == "device" or (engine is None and _device_draw_enabled()))
        if donorids is not None:  # _stats.py:20-32
            G, Y = np.asarray(donorids), np.asarray(y_std)
            Gu = np.unique(G)
            rep = np.array([np.where(G == g)[0][0] for g in Gu])
            Gind = np.searchsorted(Gu, G)
            if (Y[rep][Gind] != Y).any():
                print("ERROR: the value of Y is not identical within each group of samples")
                self._fail, self._job = True, None
                return
            self._map = (rep, Gind)
            off, pos = np.array([0, len(rep)], dtype=np.int32), None
        else:  # _stats.py:4-18
            off, pos = _batch_blocks(batches)
        if use_device:
            self._dev = self._job = _lib.DevicePermJob(off, pos, num, device)
        else:
            self._job = _lib.HostPermJob(off, pos, num)

    def done(self):
        return self._job is None or self._job.done()

    def cancel(self):
        """Joins the draw and hands numpy's generator its advanced state back.  Raises RedoWithHostDraw when
        the device engine asks for a repeat."""
        if self._dev is not None:
            if not self._dev.finish():
                raise RedoWithHostDraw()
        elif self._job is not None:
            self._job.result()

    def result(self):
        if self._fail:
            raise TypeError("'NoneType' object is not subscriptable")  # what the reference dies with
        if self._dev is not None and not self._dev.finish():
            raise RedoWithHostDraw()
        out = self._job.result()
        if self._map is not None:
            rep, Gind = self._map
            out = np.ascontiguousarray(rep[out][:, Gind], dtype=np.int32)
        return out

    def result_device(self, device):
        """The index matrix as an int32 device tensor (asynchronous)."""
        import torch
        if self._fail:
            raise TypeError("'NoneType' object is not subscriptable")
        if self._dev is not None:
            out = self._dev.result_tensor_device()
            if self._map is not None:
                rep, Gind = self._map
                rep_t = torch.as_tensor(rep.astype(np.int32), device=out.device)
                out = rep_t[out.long()][:, torch.as_tensor(Gind, device=out.device)].contiguous()
            return out
        if self._map is not None:
            return torch.from_numpy(self.result()).to(device, non_blocking=True)
        return self._job.result_tensor().to(device, non_blocking=True)


def conditional_permutation_matrix(B, num):
    """The transpose of ``conditional_permutation_indices`` as int32 [num x n] (row k = permutation
    k, the layout ``cna_perm_stats`` consumes), drawn by the native restatement of numpy's legacy
    generator: the same random stream bit for bit, with the transform, the per-column argsort and
    the scatter running on all host threads.  Advances ``np.random``'s global state exactly like
    the reference's calls."""
    from .. import _lib
    off, pos = _batch_blocks(B)
    return _lib.host_perm_blocks(off, pos, num)


def grouplevel_permutation_matrix(G, Y, num):
    """Transpose of ``grouplevel_permutation_indices`` as int32 [num x n] (or None)."""
    from .. import _lib
    G = np.asarray(G)
    Y = np.asarray(Y)
    Gu = np.unique(G)
    rep = np.array([np.where(G == g)[0][0] for g in Gu])
    Yg = Y[rep]
    Gind = np.searchsorted(Gu, G)
    if (Yg[Gind] != Y).any():
        print("ERROR: the value of Y is not identical within each group of samples")
        return None
    order = _lib.host_perm_blocks(np.array([0, len(Yg)], dtype=np.int32), None, num)  # [num x donors]
    return np.ascontiguousarray(rep[order][:, Gind], dtype=np.int32)


def threshold_edges(t, atol=1e-8, rtol=1e-5):
    """``_stats.py:51``: histogram edges for ascending thresholds t: t^2 - atol - rtol * t^2."""
    t2 = np.asarray(t, dtype=np.float64) ** 2
    return t2 - atol - rtol * t2


def tails_from_hist(hist):
    """``_stats.py:57-59``: reverse cumulative sum, tails[k, i] = #{cells: z^2 >= edge_i}."""
    return np.flip(np.cumsum(np.flip(hist, axis=-1), axis=-1), axis=-1)


def fdr_from_counts(null_hist, rank_hist, n_null=None):
    """``_stats.py:64-83`` from binned counts: fdr_i = mean_k(tails[k, i] / ranks[i]).

    ``null_hist`` is either the per-null table [n_null x T] or, with ``n_null`` given, its sum over
    the nulls [T]: the mean over k of tails[k, i] / ranks[i] is (sum_k tails[k, i]) / (n_null *
    ranks[i]) (equal up to the rounding of n_null float64 divisions, ~1e-16 relative)."""
    tails = tails_from_hist(np.asarray(null_hist, dtype=np.int64))
    ranks = tails_from_hist(np.asarray(rank_hist, dtype=np.int64))
    with np.errstate(divide="ignore", invalid="ignore"):
        if n_null is not None:
            return tails / ranks / n_null
        return (tails / ranks).mean(axis=0)
