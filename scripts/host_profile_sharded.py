"""cProfile of the host side of a sharded association() (rank 0), launched with torchrun:

    python -m torch.distributed.run --nproc-per-node N scripts/host_profile_sharded.py [C]
"""
import cProfile
import os
import pstats
import sys
import warnings

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
for _ in range(3):
    cna.tl.association(h, **kw)
torch.cuda.synchronize()
dist.barrier()
prof = cProfile.Profile()
for _ in range(10):
    torch.cuda.synchronize()
    dist.barrier()
    prof.enable()
    cna.tl.association(h, **kw)
    prof.disable()
if dist.get_rank() == 0:
    st = pstats.Stats(prof)
    st.sort_stats("cumulative").print_stats(45)
    st.sort_stats("tottime").print_stats(30)
dist.destroy_process_group()
