"""A/B of the diffusion-step kernels on the benchmark graph (config C by default):

    python scripts/spmm_tiled_bench.py [C|B] [mode ...]      modes: old, tma, cpasync

Every mode runs in this process one after the other; a kernel that traps kills the process, so the
risky ones go last (or in a call of their own).  Prints time per step and whether the tiled result is
bit-identical to cna_diffuse_step_f32.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import CONFIGS  # noqa: E402
from cna_b200 import _lib, synth  # noqa: E402
from cna_b200.tl import _graph  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C"
modes = sys.argv[2:] or ["old", "cpasync", "tma"]
N, S, k, s_steps, K = CONFIGS[cfg]
data, meta = synth.make_dataset(N, S, k, seed=0)
g = _graph.DeviceGraph(data.obsp["connectivities"])
vals, diag = g.scaled(1, torch.float32)
ld = (S + 7) // 8 * 8
torch.manual_seed(0)
src = torch.rand((g.n, ld), device="cuda")
src[:, S:] = 0
ref = torch.empty_like(src)
_lib.diffuse_step(g.indptr, g.indices, vals, diag, src, ref, S)
torch.cuda.synchronize()


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


plan = None
for mode in modes:
    if mode == "old":
        dst = torch.empty_like(src)
        print(f"old   : {timeit(lambda: _lib.diffuse_step(g.indptr, g.indices, vals, diag, src, dst, S)):.3f} ms", flush=True)
        continue
    if plan is None:
        t0 = time.perf_counter()
        plan = _graph.TilePlan(g.indptr, g.indices, vals, g.n)
        torch.cuda.synchronize()
        print(f"plan  : {plan.n_tiles} tiles, {plan.n_sources} staged rows for {plan.nnz} edges "
              f"(dedup {plan.n_sources / plan.nnz:.3f}), built in {time.perf_counter() - t0:.2f} s", flush=True)
    sm = 0 if mode == "tma" else 1
    dst = torch.full_like(src, float("nan"))
    _lib.diffuse_step_tiled(g.indptr, plan, diag, src, dst, S, stage_mode=sm)
    torch.cuda.synchronize()
    same = torch.equal(dst[:, :S], ref[:, :S])
    err = (dst[:, :S] - ref[:, :S]).abs().max().item()
    print(f"{mode:6s}: bit-identical={same} max|diff|={err:.3g}", flush=True)
    print(f"{mode:6s}: {timeit(lambda: _lib.diffuse_step_tiled(g.indptr, plan, diag, src, dst, S, stage_mode=sm)):.3f} ms", flush=True)
