"""Wall-clock stage breakdown of association() (CNA_B200_TIMING=1) at a bench configuration."""
import os
import sys
import warnings

os.environ["CNA_B200_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import cna_b200 as cna  # noqa: E402
from cna_b200 import synth  # noqa: E402
from bench import CONFIGS  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C"
N, S, k, s, K = CONFIGS[cfg]
data, meta = synth.make_dataset(N, S, k, seed=0)
kw = dict(y=meta.case, sid_name="id", batches=meta.batch, covs=meta[["age"]], nsteps=s, Nnull=K, seed=0)
warnings.simplefilter("ignore")
h = cna.tl.to_device(data)
from bench import pin_graph  # noqa: E402
undo = pin_graph(data.obsp["connectivities"])
for mode, obj in (("resident", h), ("resident", h), ("host-pinned", data), ("host-pinned", data), ("resident", h)):
    torch.cuda.synchronize()
    print("----", mode)
    cna.tl.association(obj, **kw)
if len(sys.argv) > 2:
    from threadpoolctl import threadpool_limits
    with threadpool_limits(limits=int(sys.argv[2])):
        for mode, obj in (("resident, blas threads=" + sys.argv[2], h), ("host-pinned, blas threads=" + sys.argv[2], data)):
            torch.cuda.synchronize()
            print("----", mode)
            cna.tl.association(obj, **kw)

# ---- host-side pieces in isolation (this box's cores) ----
import time  # noqa: E402

import numpy as np  # noqa: E402

from cna_b200 import _lib  # noqa: E402
from cna_b200.tl import _stats  # noqa: E402

B = np.asarray(meta.batch)
for nt in (0, 4, 8, 16):
    off, pos = _stats._batch_blocks(B)
    np.random.seed(0)
    t = time.perf_counter()
    _lib.host_perm_blocks(off, pos, K, nt)
    print(f"host_perm_blocks n_threads={nt}: {(time.perf_counter() - t) * 1e3:.1f} ms")
G = np.random.default_rng(0).normal(size=(S, 4 * S))
G = G @ G.T
from threadpoolctl import threadpool_limits  # noqa: E402
for lim in (None, 1, 4):
    for _ in range(2):
        t = time.perf_counter()
        if lim is None:
            np.linalg.svd(G)
        else:
            with threadpool_limits(limits=lim, user_api="blas"):
                np.linalg.svd(G)
        dt = (time.perf_counter() - t) * 1e3
    print(f"svd {S}x{S} blas threads={lim}: {dt:.2f} ms")
