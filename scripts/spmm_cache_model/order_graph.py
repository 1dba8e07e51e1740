"""Step 2 of the SpMM cache model: Cuthill-McKee order on the CPU (same rule as csrc/reorder.cu: breadth-first
levels from a pseudo-peripheral root, each level sorted by first parent then node id), the permuted CSR
structure as raw int32 files for lru_model.c, and the neighbour overlap of consecutive rows.

    python scripts/spmm_cache_model/order_graph.py 1000000
    gcc -O2 -o /tmp/sim/lru_model scripts/spmm_cache_model/lru_model.c
    /tmp/sim/lru_model N rows_per_cta ctas_per_sm l1_rows mode [edges_per_turn K l2_rows]
        mode 0: CTAs in grid order   1: one contiguous slice per SM   2: banded per-SM runs of K rows
    gcc -O2 -o /tmp/sim/local_greedy scripts/spmm_cache_model/local_greedy.c
    /tmp/sim/local_greedy N block window     # prototype of csrc/order_host.cpp; writes the refined graph as
                                             # size N + 1, which lru_model reads when given N + 1
Calibration: l1_rows = 128 reproduces the measured L1 hit rate of the shipped kernel (23.5 %).
"""
import sys, time, numpy as np
N=int(sys.argv[1])
z=np.load(f'/tmp/sim/graph_{N}.npz'); indptr=z['indptr'].astype(np.int64); indices=z['indices'].astype(np.int64)
n=len(indptr)-1
deg=np.diff(indptr)

def gather_edges(front):
    starts=indptr[front]; cnt=deg[front]
    tot=int(cnt.sum())
    if tot==0: return np.empty(0,np.int64), np.empty(0,np.int64)
    offs=np.repeat(starts-np.concatenate([[0],np.cumsum(cnt)[:-1]]), cnt)+np.arange(tot)
    return indices[offs], np.repeat(np.arange(len(front)), cnt)

def bfs(root, level):
    order=[np.array([root])]; level[root]=0; front=order[0]; placed=1; lvl=0
    while True:
        cols,pidx=gather_edges(front)
        m=level[cols]<0
        cols,pidx=cols[m],pidx[m]
        if len(cols)==0: break
        o=np.lexsort((pidx,cols)); cols,pidx=cols[o],pidx[o]
        first=np.concatenate([[True],cols[1:]!=cols[:-1]])
        nodes,fp=cols[first],pidx[first]
        o=np.lexsort((nodes,fp)); nodes=nodes[o]
        lvl+=1; level[nodes]=lvl
        order.append(nodes); front=nodes
    return np.concatenate(order), lvl

t=time.time()
level=np.full(n,-1,np.int64)
root=int(np.argmin(deg))
o,l=bfs(root,level)
far=o[-1]
print('first sweep levels',l,'reached',len(o),time.time()-t,flush=True)
level[:]=-1
order,l=bfs(far,level)
rest=np.nonzero(level<0)[0]
print('levels',l,'reached',len(order),'rest',len(rest),time.time()-t,flush=True)
order=np.concatenate([order,rest])
inv=np.empty(n,np.int64); inv[order]=np.arange(n)
# permuted CSR, per-row edge order kept
newdeg=deg[order]; newptr=np.concatenate([[0],np.cumsum(newdeg)])
cols,_=gather_edges(order)
newidx=inv[cols]
band=np.abs(newidx-np.repeat(np.arange(n),newdeg))
print('band pct 50/90/99/99.5', np.percentile(band,[50,90,99,99.5]))
newptr.astype(np.int32).tofile(f'/tmp/sim/ptr_{N}.bin'); newidx.astype(np.int32).tofile(f'/tmp/sim/idx_{N}.bin')
# neighbour overlap of adjacent rows
import random
ov=[]
for r in random.sample(range(1,n),2000):
    a=set(newidx[newptr[r]:newptr[r+1]]); b=set(newidx[newptr[r-1]:newptr[r]])
    ov.append(len(a&b)/max(len(a),1))
print('mean adjacent-row neighbour overlap',np.mean(ov))
ov=[]
for r in random.sample(range(8,n),2000):
    a=set(newidx[newptr[r]:newptr[r+1]]); b=set(newidx[newptr[r-7]:newptr[r]])
    ov.append(len(a&b)/max(len(a),1))
print('mean overlap with previous 7 rows',np.mean(ov))
ov=[]
for r in random.sample(range(40,n),2000):
    a=set(newidx[newptr[r]:newptr[r+1]]); b=set(newidx[newptr[r-39]:newptr[r]])
    ov.append(len(a&b)/max(len(a),1))
print('mean overlap with previous 39 rows',np.mean(ov))
