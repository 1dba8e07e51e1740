"""Wall-clock stage breakdown of a sharded association() (rank 0's marks), launched with torchrun:

    CNA_B200_TIMING=1 python -m torch.distributed.run --nproc-per-node N scripts/stage_times_sharded.py [C]
"""
import os
import sys
import warnings

os.environ["CNA_B200_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import cna_b200 as cna  # noqa: E402
from bench import CONFIGS  # noqa: E402
from cna_b200 import synth  # noqa: E402
from cna_b200.sharded import shard_to_device  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C"
N, S, k, s, K = CONFIGS[cfg]
data, meta = synth.make_dataset(N, S, k, seed=0)
kw = dict(y=meta.case, sid_name="id", batches=meta.batch, covs=meta[["age"]], nsteps=s, Nnull=K, seed=0)
warnings.simplefilter("ignore")
h = shard_to_device(data)
for mode, obj in (("resident", h), ("resident", h), ("resident", h), ("one-shot", None), ("one-shot", None)):
    torch.cuda.synchronize()
    dist.barrier()
    if dist.get_rank() == 0:
        print("----", mode, flush=True)
    cna.tl.association(obj if obj is not None else shard_to_device(data, resident=False), **kw)
import time  # noqa: E402
reps = int(os.environ.get("STAGE_REPS", "0"))
for i in range(reps):
    torch.cuda.synchronize()
    dist.barrier()
    if dist.get_rank() == 0:
        print("---- resident rep", i, flush=True)
    t0 = time.perf_counter()
    cna.tl.association(h, **kw)
    torch.cuda.synchronize()
    print(f"rank {dist.get_rank()} rep {i}: {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
if os.environ.get("STAGE_B2B"):
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for i in range(10):
        if dist.get_rank() == 0:
            print("---- back-to-back", i, flush=True)
        t1 = time.perf_counter()
        cna.tl.association(h, **kw)
        print(f"rank {dist.get_rank()} b2b {i}: {1e3 * (time.perf_counter() - t1):.2f} ms", flush=True)
    torch.cuda.synchronize()
    print(f"rank {dist.get_rank()} back-to-back mean: {1e2 * (time.perf_counter() - t0):.2f} ms", flush=True)
dist.destroy_process_group()
