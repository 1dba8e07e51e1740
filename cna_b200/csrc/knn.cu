// Exact brute-force kNN behind cna_b200.pp.neighbors (the graph scanpy.pp.neighbors would build, read by the
// reference at _nam.py:12-19) and the synthetic-data generator; it is not part of the timed path.  One thread per
// query, candidate tiles broadcast from shared memory, a sorted top-k list per thread.
#include "common.cuh"

namespace cna {

constexpr int kKnnTileFloats = 8192;  // 32 KB of candidates per tile
constexpr int kKnnMaxK = 64;

template <int DIM>
__global__ void __launch_bounds__(128)
knn_kernel(const float *__restrict__ pts, int64_t n, int k, int64_t q0, int64_t nq, int32_t *__restrict__ idx,
           float *__restrict__ dist2) {
    constexpr int kKnnTile = kKnnTileFloats / DIM;
    __shared__ float tile[kKnnTile * DIM];
    const int64_t ql = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;  // query index inside [q0, q0 + nq)
    const int64_t qi = q0 + ql;
    float q[DIM];
    bool live = ql < nq;
#pragma unroll
    for (int d = 0; d < DIM; ++d) q[d] = live ? pts[qi * DIM + d] : 0.f;
    float bd[kKnnMaxK];
    int bi[kKnnMaxK];
    for (int t = 0; t < k; ++t) {
        bd[t] = INFINITY;
        bi[t] = -1;
    }
    float worst = INFINITY;
    for (int64_t c0 = 0; c0 < n; c0 += kKnnTile) {
        int cnt = int(n - c0 < kKnnTile ? n - c0 : kKnnTile);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt * DIM; t += blockDim.x) tile[t] = pts[c0 * DIM + t];
        __syncthreads();
        if (!live) continue;
        for (int c = 0; c < cnt; ++c) {
            float d2 = 0.f;
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                float df = q[d] - tile[c * DIM + d];
                d2 = fmaf(df, df, d2);
            }
            if (d2 < worst && c0 + c != qi) {
                int pos = k - 1;
                while (pos > 0 && bd[pos - 1] > d2) {
                    bd[pos] = bd[pos - 1];
                    bi[pos] = bi[pos - 1];
                    --pos;
                }
                bd[pos] = d2;
                bi[pos] = int(c0 + c);
                worst = bd[k - 1];
            }
        }
    }
    if (live)
        for (int t = 0; t < k; ++t) {
            idx[ql * k + t] = bi[t];
            dist2[ql * k + t] = bd[t];
        }
}

}  // namespace cna

using namespace cna;

extern "C" int cna_knn_bruteforce_range(const float *points, int64_t n, int dim, int k, int64_t q0, int64_t nq,
                                        int32_t *idx, float *dist2, void *stream) {
    CNA_REQUIRE(n > 0 && k > 0 && k <= kKnnMaxK && k < n, "cna_knn_bruteforce: need 0 < k <= 64, k < n");
    CNA_REQUIRE(q0 >= 0 && nq >= 0 && q0 + nq <= n, "cna_knn_bruteforce: bad query range");
    if (nq == 0) return CNA_OK;
    unsigned grid = unsigned((nq + 127) / 128);
    cudaStream_t st = as_stream(stream);
    switch (dim) {
        case 4: knn_kernel<4><<<grid, 128, 0, st>>>(points, n, k, q0, nq, idx, dist2); break;
        case 8: knn_kernel<8><<<grid, 128, 0, st>>>(points, n, k, q0, nq, idx, dist2); break;
        case 16: knn_kernel<16><<<grid, 128, 0, st>>>(points, n, k, q0, nq, idx, dist2); break;
        case 32: knn_kernel<32><<<grid, 128, 0, st>>>(points, n, k, q0, nq, idx, dist2); break;
        case 64: knn_kernel<64><<<grid, 128, 0, st>>>(points, n, k, q0, nq, idx, dist2); break;
        default: return set_error(CNA_ERR_INVALID, "cna_knn_bruteforce: dim must be 4, 8, 16, 32 or 64 (pad with zeros), got %d", dim);
    }
    CNA_LAUNCHED("knn_kernel");
    return CNA_OK;
}

extern "C" int cna_knn_bruteforce(const float *points, int64_t n, int dim, int k, int32_t *idx,
                                  float *dist2, void *stream) {
    return cna_knn_bruteforce_range(points, n, dim, k, 0, n, idx, dist2, stream);
}
