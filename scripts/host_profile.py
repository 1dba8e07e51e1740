"""cProfile of one association() call at a bench configuration (host-side hot spots)."""
import cProfile
import io
import os
import pstats
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import cna_b200 as cna  # noqa: E402
from cna_b200 import synth  # noqa: E402
from bench import CONFIGS  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C"
resident = len(sys.argv) <= 2 or sys.argv[2] != "host"
N, S, k, s, K = CONFIGS[cfg]
data, meta = synth.make_dataset(N, S, k, seed=0)
kw = dict(y=meta.case, sid_name="id", batches=meta.batch, covs=meta[["age"]], nsteps=s, Nnull=K, seed=0)
h = cna.tl.to_device(data) if resident else data
warnings.simplefilter("ignore")
for _ in range(2):
    cna.tl.association(h, **kw)
torch.cuda.synchronize()
pr = cProfile.Profile()
reps = 20
for _ in range(reps):
    torch.cuda.synchronize()
    pr.enable()
    cna.tl.association(h, **kw)
    pr.disable()
out = io.StringIO()
st = pstats.Stats(pr, stream=out)
print(f"{reps} calls profiled; divide by {reps}")
st.sort_stats("cumulative").print_stats(60)
st.sort_stats("tottime").print_stats(45)
print(out.getvalue())
