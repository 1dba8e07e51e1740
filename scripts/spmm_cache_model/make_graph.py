"""Step 1 of the SpMM cache model: build the benchmark graph on the CPU (k-d tree kNN
instead of the GPU brute-force kernel; same generator, same seed) and save its CSR structure.

    python scripts/spmm_cache_model/make_graph.py 1000000      # ~4 min, writes /tmp/sim/graph_<N>.npz
"""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from cna_b200 import synth
t=time.time()
N=int(sys.argv[1])
data, meta = synth.make_dataset(N, 200, 30, seed=0, knn="cpu", dim=6)
A = data.obsp['connectivities'].tocsr()
print('gen', time.time()-t, A.nnz, flush=True)
os.makedirs('/tmp/sim', exist_ok=True)
np.savez(f'/tmp/sim/graph_{N}.npz', indptr=A.indptr, indices=A.indices)
