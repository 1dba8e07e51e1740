#!/usr/bin/env python
"""Benchmark of the CNA hot path: cells/sec through nam() + association().

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C|B|A]

One "step" is one ``cna.tl.association`` call (which builds the NAM internally) on a synthetic
AnnData of the named shape.  Default workload = BASELINE.json configs[2] ("C"): 1M cells, 200
samples, k=30, s=3 diffusion steps, 10 000 permutations.

  value : whole-job cells/sec with the kNN graph and sample codes already resident in HBM
          (``cna.tl.to_device``); everything else — host RNG for the permutations, the n x n SVD,
          the F tests, result read-back into ``data.obs`` — is inside the timed region.
  e2e   : the same call on the host AnnData (scipy CSR in pinned host memory): H2D of the graph and
          D2H of the per-cell results happen inside the timed region, every step.
  roofline / rooflines : per-kernel achieved bandwidth / flop rate from CUDA-event pairs recorded
          around every library call during the timed steps, against MEASURED_PEAKS.json.
  cpu_baseline : the oracle in ``faithful`` mode (the reference's own cost structure: scipy SpMM,
          per-step kurtosis, Python loop over permutations, np.histogram per null, per-cell
          Series.apply) on a bounded sample of the same workload, on this box's host cores.

``--impl reference`` times only that CPU arm.  Timing: CUDA events on the launching stream, barrier
and synchronize on both sides, max over ranks.  Inputs (0.45 GB graph, 0.8 GB state) are far larger
than the 126 MB L2, so no explicit L2 flush is needed between iterations.
"""
import argparse
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONFIGS = {
    # name: (cells, samples, k, diffusion steps, permutations)
    "A": (10_000, 50, 15, 3, 1000),
    "B": (100_000, 100, 15, 3, 1000),
    "C": (1_000_000, 200, 30, 3, 10_000),
    "E": (10_000_000, 500, 30, 4, 10_000),
    "T": (200_000, 60, 15, 4, 500),  # smoke test of the row-block (config E) path at a size that takes seconds
    "E1": (1_250_000, 500, 30, 4, 10_000),  # one rank's share of config E on one GPU (its own graph: no halo)
}
METRIC = "cells/sec through nam()+association(), 1M cells/200 samples/10k perms"
# CPU sample: same samples / k / steps, cells and permutations scaled down by the same factor
CPU_SAMPLE_SCALE = 20
# DRAM traffic of one SpMM launch at config C from the committed `ncu --set full` captures
# (profiles/r02c_ncu_full_summary.txt: 2.605 GB read + 0.788 GB written; profiles/r01c_ncu_spmm_summary.txt:
# 2.600 + 0.780 GB for the plain step; 31.6 GB before the cell reordering)
SPMM_DRAM_TRAFFIC_GB = 3.393


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-full", action="store_true",
                    help="also run the CPU oracle once on the SAME full-size AnnData (same-config baseline + result check; minutes)")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_name(cfg):
    n, S, k, s, K = CONFIGS[cfg]
    return f"{n} cells, {S} samples, k={k}, s={s}, {K} permutations (config {cfg})"


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle with the reference's cost structure on a bounded sample
# ---------------------------------------------------------------------------------------------
def cpu_sample_spec(cfg):
    n, S, k, s, K = CONFIGS[cfg]
    scale = CPU_SAMPLE_SCALE if n >= 100_000 else 1
    return max(n // scale, 20 * S // 2), S, k, s, max(K // scale, 50)


def make_cpu_sample(cfg):
    from cna_b200 import synth
    n, S, k, s, K = cpu_sample_spec(cfg)
    data, meta = synth.make_dataset(n, S, k, seed=0, dim=6 if CONFIGS[cfg][0] >= 1_000_000 else None,
                                    knn="cpu", device="cpu")
    kw = dict(y=meta.case, sid_name="id", batches=meta.batch, covs=meta[["age"]], nsteps=s, Nnull=K, seed=0)
    return data, kw, f"{n} cells, {S} samples, k={k}, s={s}, {K} permutations (cells and permutations = config {cfg} / {CONFIGS[cfg][0] // n})"


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def time_cpu_arm(data, kw, steps, warmup):
    from oracle import cna_oracle as orc
    times = []
    for i in range(warmup + steps):
        d = type(data)(data.obs.copy(), data.obsp["connectivities"])
        t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            orc.association(d, faithful=True, **kw)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times


def oracle_full(data, kw):
    """The CPU oracle (reference cost structure) once on the very AnnData the GPU arm ran on: a same-config
    baseline and the result fingerprint the GPU `check` block is compared with."""
    from oracle import cna_oracle as orc
    d = type(data)(data.obs[["id"]].copy(), data.obsp["connectivities"])
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = orc.association(d, faithful=True, return_full=True, **kw)
    dt = time.perf_counter() - t0
    coef, fdr = d.obs["coef"].to_numpy(), d.obs["coef_fdr"].to_numpy()
    check = {
        "p": float(res.p), "k": int(res.k), "n_kept": int(res.kept.sum()),
        "n_fdr05": int((fdr <= 0.05).sum()), "n_fdr10": int((fdr <= 0.10).sum()),
        "fdr_5p_t": None if res.fdr_5p_t is None else float(res.fdr_5p_t),
        "fdr_10p_t": None if res.fdr_10p_t is None else float(res.fdr_10p_t),
        "svs": [float(v) for v in res.namresid_svs.to_numpy()[:4]], "sum_abs_coef": float(np.nansum(np.abs(coef))),
    }
    return dt, check


def run_reference(args):
    """--impl reference: the CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    data, kw, sample = make_cpu_sample(args.config)
    times = time_cpu_arm(data, kw, args.steps, min(args.warmup, 1))
    n_cells = len(data.obs)
    value = n_cells * len(times) / sum(times)
    cores = cpu_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": workload_name(args.config), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        # the CPU arm runs a 1 / sample_scale sample of the workload (cells and permutations): its cells/s is
        # an upper bound for the full-size CPU run (profiles/r02_oracle_C.json has that one: same-config)
        "sample_scale": CONFIGS[args.config][0] // n_cells, "same_config": CONFIGS[args.config][0] == n_cells,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region, in-process through NVML (a thread that
    wakes every 20 ms; starting an `nvidia-smi -lms` child instead costs ~100 ms of driver-wide locking
    at start-up, which showed up as +4 ms per step on a 2-GPU run whose timed region is only 0.1 s)."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index = index
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self.stop = None
        self.thread = None

    def __enter__(self):
        try:
            import threading

            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                phys = int(vis.split(",")[self.index])
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.stop = threading.Event()

            def loop():
                while not self.stop.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        for name, bit in self.BAD.items():
                            if mask & bit:
                                self.reasons.add(name)
                    except Exception:
                        pass
                    self.stop.wait(0.02)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
        return self

    def __exit__(self, *exc):
        if self.thread is not None:
            self.stop.set()
            self.thread.join(timeout=2)

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        if self.sm:
            out.update(sm_mhz=float(np.median(self.sm)), samples=len(self.sm), source="nvml")
        return out


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tensor=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    tensor_burst=p["bf16_tflops"], source="measured")
    return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1590.0, source="fallback")


def pin_graph(A):
    """Register the scipy CSR buffers as pinned host memory (the e2e contract: inputs start in
    pinned host memory).  Returns a callable that unregisters them."""
    import torch
    rt = torch.cuda.cudart()
    regs = []
    for arr in (A.data, A.indices, A.indptr):
        if int(rt.cudaHostRegister(arr.ctypes.data, arr.nbytes, 0)) == 0:
            regs.append(arr)

    def undo():
        for arr in regs:
            rt.cudaHostUnregister(arr.ctypes.data)
    return undo


def check_block(obs, key="coef"):
    """Result fingerprint of the last association() call (every rank of a sharded run ends with the full
    data.obs columns): discriminates a correct run from a fast wrong one, and is compared across N and
    against the CPU oracle (profiles/r02_oracle_C.json)."""
    from cna_b200.tl import _association as A
    coef = obs[key].to_numpy()
    fdr = obs[key + "_fdr"].to_numpy() if key + "_fdr" in obs else np.full(len(coef), np.nan)
    last = A.LAST
    return {
        "p": float(last.p), "k": int(last.k), "n_kept": int(np.isfinite(coef).sum()),
        "n_fdr05": int((fdr <= 0.05).sum()), "n_fdr10": int((fdr <= 0.10).sum()),
        "fdr_5p_t": None if last.fdr_5p_t is None else float(last.fdr_5p_t),
        "fdr_10p_t": None if last.fdr_10p_t is None else float(last.fdr_10p_t),
        "svs": [float(v) for v in last.svs[:4]], "sum_abs_coef": float(np.nansum(np.abs(coef))),
    }


def checks_agree(a, b, rtol=1e-5):
    """Integer fields exactly, floats to the north-star tolerance (the Gram is summed in a different
    order on a different number of GPUs; cells within rounding of an FDR threshold may flip)."""
    for key in ("p", "k", "n_kept"):
        if a[key] != b[key]:
            return False
    for key in ("n_fdr05", "n_fdr10"):
        if abs(a[key] - b[key]) > max(3, 1e-4 * max(a[key], b[key])):
            return False
    for key in ("fdr_5p_t", "fdr_10p_t", "sum_abs_coef"):
        if (a[key] is None) != (b[key] is None) or (a[key] is not None and abs(a[key] - b[key]) > rtol * abs(b[key])):
            return False
    return bool(np.allclose(a["svs"], b["svs"], rtol=rtol))


def rooflines(prof, steps, N, S, n, nnz, K, Kl, s_steps, peaks):
    """Per-kernel achieved rates from the CUDA-event profile of the timed steps.
    Algorithmic bytes / flops per launch follow SURVEY.md 8(d) and DESIGN.md."""
    b_spmm = 8 * nnz + 4 * (N + 1) + 4 * N + 2 * 4 * N * S
    spec = {
        "cna_diffuse_step_f32": ("hbm", b_spmm, "diffusion SpMM (one step)"),
        "cna_diffuse_step_f32_qc": ("hbm", b_spmm, "diffusion SpMM, last step, with the QC batch-kurtosis fused in"),
        "cna_diffuse_onehot": ("hbm", b_spmm, "diffusion step 1 from the one-hot indicator"),
        "cna_resid_pass": ("hbm", 2 * 4 * N * S, "select/centre/residualise/standardise/ncorr pass"),
        "cna_gram": ("tensor", 2.0 * n * n * N, "Gram X^T X (CUDA cores)"),
        "cna_gram_tc": ("tensor", 2.0 * n * n * N, "Gram X^T X (tcgen05, fp16 hi/lo split, fp64 flush)"),
        "cna_null_hist": ("tensor", 2.0 * N * n * Kl, "null GEMM + threshold histogram (CUDA cores)"),
        "cna_null_hist_tc": ("tensor", 2.0 * N * n * Kl, "null GEMM + threshold histogram (tcgen05, fp16 hi/lo split)"),
        "cna_null_hist_tc_dev": ("tensor", 2.0 * N * n * Kl, "null GEMM + threshold histogram (tcgen05, fp16 hi/lo split; thresholds read from device memory)"),
        "cna_perm_stats": (None, None, "permutation engine (fp64)"),
    }
    total = sum(ms for _, ms in prof.values()) or 1.0
    out = []
    for name, (calls, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        bound, work, what = spec.get(name, (None, None, name))
        avg = ms / max(calls, 1)
        item = {"kernel": name, "what": what, "launches_per_step": calls / steps, "avg_ms": avg,
                "share_of_device_time": ms / total}
        if bound == "hbm":
            ach = work / (avg * 1e-3) / 1e9
            item.update(bound="hbm", achieved=ach, peak=peaks["hbm"], unit="GB/s", frac=ach / peaks["hbm"],
                        algorithmic_bytes=work)
        elif bound == "tensor":
            ach = work / (avg * 1e-3) / 1e12
            item.update(bound="tensor", achieved=ach, peak=peaks["tensor"], unit="TFLOP/s",
                        frac=ach / peaks["tensor"], algorithmic_flops=work)
        out.append(item)
    return out


def run_ours(args):
    # stdout carries exactly one JSON line: anything libraries print there (NCCL's version banner,
    # warnings) goes to stderr instead
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import cna_b200 as cna
    from cna_b200 import _lib, synth
    _lib.load()

    N, S, k, s_steps, K = CONFIGS[args.config]
    t0 = time.perf_counter()
    comm = None
    if world > 1:
        from cna_b200.sharded import Comm
        comm = Comm()
    # the kNN search of the generator is split over the ranks; at config E every rank also keeps only its
    # own block of rows of the graph on the host (10M x 10M, 4e8 stored edges)
    blocks = args.config in ("E", "T") and world > 1
    data, meta = synth.make_dataset(N, S, k, seed=0, dim=6 if N >= 1_000_000 else None, comm=comm, row_block=blocks)
    gen_s = time.perf_counter() - t0
    A = data.obsp["connectivities"]
    nnz = int(A.nnz)
    host_graph_bytes = A.data.nbytes + A.indices.nbytes + A.indptr.nbytes
    if blocks:
        t = torch.tensor([nnz, host_graph_bytes], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        nnz, host_graph_bytes = int(t[0].item()), int(t[1].item())
    kw = dict(y=meta.case, sid_name="id", batches=meta.batch, covs=meta[["age"]], nsteps=s_steps, Nnull=K, seed=0)
    n = S
    Kl = min(1000, K)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    if world > 1:
        from cna_b200.sharded import shard_to_device
        handle = shard_to_device(data)
        step_dev = lambda: cna.tl.association(handle, **kw)  # noqa: E731
        step_e2e = lambda: cna.tl.association(shard_to_device(data, resident=False), **kw)  # noqa: E731
    else:
        handle = cna.tl.to_device(data)
        step_dev = lambda: cna.tl.association(handle, **kw)  # noqa: E731
        step_e2e = lambda: cna.tl.association(data, **kw)  # noqa: E731

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(args.warmup):
            step_dev()
        launches0 = _lib.launch_count()
        with ClockSampler(local) as clk:
            ms = timed(step_dev, args.steps)
        launches = _lib.launch_count() - launches0
        # per-kernel CUDA-event pairs are recorded in a separate, untimed pass of the same step
        prof_steps = min(args.steps, 5)
        _lib.profile_start()
        for _ in range(prof_steps):
            p_value = step_dev()
        prof = _lib.profile_stop()
        check = check_block(data.obs)
        # standalone cna.tl.nam() (NAM + QC, _nam.py:179) on the resident graph; the samples x cells DataFrame it
        # returns is 1.6 GB at config C and is copied to the host inside this number
        nam_ms = None
        if world == 1 and args.config in ("A", "B", "C"):
            nam_call = lambda: cna.tl.nam(handle, "id", batches=meta.batch, nsteps=s_steps)  # noqa: E731
            nam_call()
            nam_ms = timed(nam_call, 3) / 3
        e2e = None
        if not args.no_e2e:
            # the same call on the caller's own (pageable) scipy buffers: what a user gets without registering
            # the CSR arrays as pinned memory first (fewer steps: the number is for the record, not the headline)
            pg_steps = max(3, args.steps // 4)
            step_e2e()
            ms_pageable = timed(step_e2e, pg_steps) / pg_steps
            undo = pin_graph(A)
            for _ in range(min(args.warmup, 2)):
                step_e2e()
            ms_e2e = timed(step_e2e, args.steps)
            undo()
            check_e2e = check_block(data.obs)
            check["e2e_agrees"] = checks_agree(check_e2e, check)
            h2d = host_graph_bytes + 4 * N * world  # summed over the ranks (each uploads its block of rows)
            d2h = 2 * 8 * N + 8 * n * n + 8 * K * 5
            e2e = {"value": N * args.steps / (ms_e2e * 1e-3), "unit": "cells/s", "ms_per_step": ms_e2e / args.steps,
                   "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "input_memory": "scipy CSR buffers registered as pinned host memory (cudaHostRegister) outside the timed region",
                   "pageable": {"value": N / (ms_pageable * 1e-3), "unit": "cells/s", "ms_per_step": ms_pageable,
                                "steps": pg_steps, "input_memory": "the caller's pageable scipy CSR buffers as they are"}}

    if rank != 0:
        dist.destroy_process_group()
        return
    if world > 1 and not blocks:  # the same call on one GPU (outside every timed region): the shards must reproduce it
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            cna.tl.association(cna.tl.to_device(data), **kw)
        single = check_block(data.obs)
        check["single_gpu"] = single
        check["agrees_with_single_gpu"] = checks_agree(check, single)
        assert check["agrees_with_single_gpu"], (check, single)
    peaks = measured_peaks()
    # per-kernel work of one launch = this rank's share of the cells (and of the stored edges)
    roofs = rooflines(prof, prof_steps, N // world, S, n, nnz // world, K, Kl, s_steps, peaks)
    spmm = next((r for r in roofs if r["kernel"] == "cna_diffuse_step_f32"), None)
    primary = None
    if spmm:
        primary = {"kernel": "cna_diffuse_step_f32 (CSR SpMM diffusion step)", "bound": "hbm",
                   "achieved": spmm["achieved"], "peak": spmm["peak"], "unit": "GB/s", "frac": spmm["frac"],
                   "traffic": SPMM_DRAM_TRAFFIC_GB if (args.config == "C" and world == 1) else None,
                   "traffic_unit": "GB per launch (dram__bytes_read.sum + dram__bytes_write.sum of one spmm_f32_kernel "
                                   "launch, profiles/r02c_ncu_full_summary.txt)",
                   "algorithmic_gb": spmm["algorithmic_bytes"] / 1e9,
                   "peak_source": peaks["source"] + " (MEASURED_PEAKS.json hbm_gbs)"}
    line = {
        "metric": METRIC, "value": N * args.steps / (ms * 1e-3), "unit": "cells/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32 state, fp16x2-split tensor-core GEMMs with f32 accumulate, f64 statistics", "data": "synthetic",
        "config": {"workload": workload_name(args.config), "nnz": nnz, "l2": "inputs larger than L2 (no flush)",
                   "parallelism": f"cell-axis shards x{world}" if world > 1 else "single GPU",
                   "p_value": p_value, "datagen_s": round(gen_s, 1)},
        "clocks": clk.summary(), "e2e": e2e, "gpu_launches": launches / args.steps, "check": check,
        "nam_standalone_ms": nam_ms,
        "roofline": primary, "rooflines": roofs,
    }
    if not args.no_cpu_baseline and world == 1 and args.config not in ("E", "T"):
        cdata, ckw, sample = make_cpu_sample(args.config)
        t = time_cpu_arm(cdata, ckw, 1, 0)
        line["cpu_baseline"] = {"value": len(cdata.obs) / t[0], "unit": "cells/s", "cores": cpu_threads(),
                                "kind": "port", "sample": sample, "seconds": t[0]}
    if args.cpu_full and world == 1:
        secs, ocheck = oracle_full(data, kw)
        line["cpu_baseline_full"] = {"value": N / secs, "unit": "cells/s", "seconds": secs, "cores": cpu_threads(),
                                     "kind": "port", "same_config": True,
                                     "sample": workload_name(args.config) + ", the same AnnData as the GPU arm",
                                     "check": ocheck, "gpu_check_agrees": checks_agree(check, ocheck)}
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
