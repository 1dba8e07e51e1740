"""Micro-benchmarks of single kernels at config-C shapes on synthetic inputs (CUDA events, 20 launches
after 3 warm-ups, inputs larger than L2).  Used to A/B kernel variants on the GPU box:

    python scripts/kernel_bench.py resid median          # any subset of: resid median
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402
import torch  # noqa: E402

from cna_b200 import _lib  # noqa: E402
from cna_b200.tl import _nam  # noqa: E402


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def bench_resid(N=1_000_000, S=200, nb=4, ncov=1):
    rng = np.random.default_rng(0)
    counts = rng.integers(4000, 6000, S).astype(float)
    s = torch.rand((N, S), device="cuda") * 3
    st = _nam.NamState(s, S, pd.Index(np.arange(S)), counts, pd.RangeIndex(N))
    batches = np.arange(S) % nb
    covs = rng.normal(size=(S, ncov))
    y = rng.normal(size=S)
    y = (y - y.mean()) / y.std()
    colmap = np.arange(S, dtype=np.int32)
    ms = timeit(lambda: _nam.resid_nam_device(st, colmap, covs, batches, y, ridges=[1e5], want_x=False, speculate=True))
    print(f"resid pass (+ median of the ridge walk): {ms:.3f} ms per call  [CNA_RESID_R={os.environ.get('CNA_RESID_R')}]")
    prof_only(lambda: _nam.resid_nam_device(st, colmap, covs, batches, y, ridges=[1e5], want_x=False, speculate=True))


def prof_only(fn):
    _lib.profile_start()
    for _ in range(5):
        fn()
    for k, (c, ms) in _lib.profile_stop().items():
        print(f"    {k}: {ms / c:.3f} ms")


def bench_median(N=1_000_000):
    v = torch.rand(N, dtype=torch.float64, device="cuda") * 7
    out = torch.empty(2, dtype=torch.float64, device="cuda")
    print(f"median of {N}: {timeit(lambda: _lib.median(v, None, out)):.3f} ms")
    valid = (torch.rand(N, device="cuda") > 0.1).to(torch.uint8)
    print(f"median of {N}, masked: {timeit(lambda: _lib.median(v, valid, out)):.3f} ms")


if __name__ == "__main__":
    what = sys.argv[1:] or ["resid", "median"]
    if "resid" in what:
        bench_resid()
    if "median" in what:
        bench_median()
