// Locality-restoring cell order for the diffusion SpMM (SURVEY.md section 7 "hard parts").
//
// Real AnnData objects store cells sample by sample, so a cell's kNN neighbours are scattered over
// the whole state matrix: every gathered 800-byte row of `s` comes from HBM (measured: 31.6 GB of
// DRAM traffic per step against 1.9 GB of compulsory bytes).  The reference has no counterpart: this
// is a property of how the graph is laid out in HBM.  A Cuthill-McKee ordering (breadth-first levels,
// each level sorted by the position of its first parent) puts 99.5 % of the edges of the 1M-cell
// graph within +-78 000 rows, i.e. inside the L2 working set of the SMs that are active at the same
// time.  The ordering only renames cells: rows keep their edges in their original order, so every
// floating-point sum is performed in exactly the same order as without it.
//
// Device part: one frontier expansion per BFS level (first parent via atomicMin, discovery via
// atomicCAS) and the permuted copy of the CSR.  The host (tl/_graph.py) sorts each level with
// torch.sort and drives the loop.
#include "common.cuh"

namespace cna {

// One warp per frontier node u (at position pos_base + i of the final order): every unvisited
// neighbour v is claimed for the next level and remembers the smallest parent position.
__global__ void bfs_expand_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                  const int32_t *__restrict__ frontier, int n_front, int pos_base,
                                  int next_level, int32_t *level, int32_t *first_parent,
                                  int32_t *next, int32_t *next_count) {
    int64_t w = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (w >= n_front) return;
    int u = frontier[w];
    int p = pos_base + int(w);
    for (int e = indptr[u] + lane, e1 = indptr[u + 1]; e < e1; e += 32) {
        int v = indices[e];
        int lv = level[v];
        if (lv >= 0 && lv != next_level) continue;  // settled in an earlier level
        if (lv < 0) {
            int old = atomicCAS(level + v, -1, next_level);
            if (old == -1) next[atomicAdd(next_count, 1)] = v;
            else if (old != next_level) continue;
        }
        atomicMin(first_parent + v, p);
    }
}

// Sort keys of the nodes discovered for the next level: (position of the first parent, node id).
__global__ void bfs_keys_kernel(const int32_t *__restrict__ next, int n, const int32_t *__restrict__ first_parent,
                                int64_t *__restrict__ keys) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int v = next[i];
        keys[i] = (int64_t(first_parent[v]) << 32) | int64_t(uint32_t(v));
    }
}

// The sorted level becomes the next stretch of the order and the next frontier.
__global__ void bfs_place_kernel(const int64_t *__restrict__ sorted_keys, int n, int64_t *__restrict__ order_out,
                                 int32_t *__restrict__ frontier) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int v = int(sorted_keys[i] & 0xFFFFFFFFll);
        order_out[i] = v;
        frontier[i] = v;
    }
}

// Row i of the new graph is row order[i] of the old one, column ids renamed through inv.
template <typename T>
__global__ void permute_csr_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                   const T *__restrict__ data, const int64_t *__restrict__ order,
                                   const int32_t *__restrict__ inv, const int32_t *__restrict__ new_indptr,
                                   int64_t n_rows, int32_t *__restrict__ new_indices, T *__restrict__ new_data) {
    int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    int64_t old = order[row];
    int e0 = indptr[old], cnt = indptr[old + 1] - e0, o0 = new_indptr[row];
    for (int t = lane; t < cnt; t += 32) {
        new_indices[o0 + t] = inv[indices[e0 + t]];
        new_data[o0 + t] = data[e0 + t];
    }
}

}  // namespace cna

using namespace cna;

extern "C" {

int cna_bfs_expand(const int32_t *indptr, const int32_t *indices, const int32_t *frontier, int n_front,
                   int pos_base, int next_level, int32_t *level, int32_t *first_parent, int32_t *next,
                   int32_t *next_count, void *stream) {
    CNA_REQUIRE(indptr && indices && frontier && level && first_parent && next && next_count && n_front >= 0,
                "cna_bfs_expand: bad arguments");
    if (n_front == 0) return CNA_OK;
    unsigned grid = unsigned((int64_t(n_front) * 32 + 255) / 256);
    bfs_expand_kernel<<<grid, 256, 0, as_stream(stream)>>>(indptr, indices, frontier, n_front, pos_base,
                                                          next_level, level, first_parent, next, next_count);
    CNA_LAUNCHED("bfs_expand_kernel");
    return CNA_OK;
}

int cna_bfs_keys(const int32_t *next, int n, const int32_t *first_parent, int64_t *keys, void *stream) {
    CNA_REQUIRE(next && first_parent && keys && n >= 0, "cna_bfs_keys: bad arguments");
    if (n == 0) return CNA_OK;
    bfs_keys_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(next, n, first_parent, keys);
    CNA_LAUNCHED("bfs_keys_kernel");
    return CNA_OK;
}

int cna_bfs_place(const int64_t *sorted_keys, int n, int64_t *order_out, int32_t *frontier, void *stream) {
    CNA_REQUIRE(sorted_keys && order_out && frontier && n >= 0, "cna_bfs_place: bad arguments");
    if (n == 0) return CNA_OK;
    bfs_place_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(sorted_keys, n, order_out, frontier);
    CNA_LAUNCHED("bfs_place_kernel");
    return CNA_OK;
}

int cna_permute_csr(const int32_t *indptr, const int32_t *indices, const void *data, int is_f64,
                    const int64_t *order, const int32_t *inv, const int32_t *new_indptr, int64_t n_rows,
                    int32_t *new_indices, void *new_data, void *stream) {
    CNA_REQUIRE(indptr && indices && data && order && inv && new_indptr && new_indices && new_data && n_rows >= 0,
                "cna_permute_csr: bad arguments");
    if (n_rows == 0) return CNA_OK;
    unsigned grid = unsigned((n_rows * 32 + 255) / 256);
    cudaStream_t st = as_stream(stream);
    if (is_f64)
        permute_csr_kernel<double><<<grid, 256, 0, st>>>(indptr, indices, static_cast<const double *>(data), order, inv,
                                                        new_indptr, n_rows, new_indices, static_cast<double *>(new_data));
    else
        permute_csr_kernel<float><<<grid, 256, 0, st>>>(indptr, indices, static_cast<const float *>(data), order, inv,
                                                       new_indptr, n_rows, new_indices, static_cast<float *>(new_data));
    CNA_LAUNCHED("permute_csr_kernel");
    return CNA_OK;
}

}  // extern "C"
