// Kernel (i): random-walk diffusion of the cells x samples state over the kNN graph.
//
// Reference: src/cna/tools/_nam.py:21-34 (diffuse_stepwise).  One step is
//     s <- A.(s / colsums[:,None]) + w * s / colsums[:,None],   colsums = A.sum(axis=0) + w
// i.e. s <- (A + wI) D^-1 s.  The D^-1 scaling is folded into the edge values once
// (cna_graph_scale), so a step is a plain CSR SpMM plus a diagonal term.
//
// Data layout: state row-major [n_rows x ld] fp32, ld a multiple of 8 floats so that every row is
// a whole number of 32-byte sectors.  A warp owns one output row; lanes own float4 column groups,
// the (index, value) pairs of the row are read coalesced 32 at a time and broadcast with shuffles,
// and four gathered source rows are kept in flight per lane.  Accumulation is in CSR order, the
// order scipy's csr_matvecs uses.  The graph is stored in a Cuthill-McKee cell order (reorder.cu), so
// the gathered rows come from L2/L1 rather than HBM; the kernel is then bound by the L1/TEX data pipe
// (profiles/): all 4.nnz.S gathered bytes pass through it whatever the hit rate.
#include "common.cuh"

namespace cna {

// ---------------------------------------------------------------------------------------------
// graph preparation
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void colsum_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                              const T *__restrict__ data, int64_t n_rows, double *colsum) {
    // one warp per row keeps the index/value reads coalesced; fp64 atomics land in L2
    int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    int e0 = indptr[row], e1 = indptr[row + 1];
    for (int e = e0 + lane; e < e1; e += 32) atomicAdd(colsum + indices[e], double(data[e]));
}

template <typename T, typename O>
__global__ void scale_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                             const T *__restrict__ data, int64_t n_rows,
                             const double *__restrict__ colsum, double w, O *__restrict__ vals,
                             O *__restrict__ diag, int64_t row_offset) {
    int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    int e0 = indptr[row], e1 = indptr[row + 1];
    for (int e = e0 + lane; e < e1; e += 32)
        vals[e] = O(double(data[e]) / (colsum[indices[e]] + w));
    if (lane == 0) diag[row] = O(w / (colsum[row + row_offset] + w));
}

// ---------------------------------------------------------------------------------------------
// first step from the one-hot indicator
// ---------------------------------------------------------------------------------------------
// The output row is assembled in a per-warp shared-memory buffer: each lane owns one edge of the
// row, lanes whose edges hit the same sample column are grouped with match.any and every member of a
// group replays the group's additions in lane (= CSR) order, so the sums are performed exactly in
// the order scipy's csr_matvecs uses and the result is deterministic.  ~150 instructions per row
// instead of ~600 for a lanes-own-columns loop over the edges.
__global__ void __launch_bounds__(256)
onehot_step_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                   const float *__restrict__ vals, const float *__restrict__ diag,
                   const int32_t *__restrict__ code, int64_t n_rows, float *__restrict__ out,
                   int64_t ld, int64_t row_offset) {
    extern __shared__ float onehot_buf[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *buf = onehot_buf + warp * ld;
    for (int64_t row = int64_t(blockIdx.x) * 8 + warp; row < n_rows; row += int64_t(gridDim.x) * 8) {
        for (int c = lane; c < ld; c += 32) buf[c] = 0.f;  // padding columns stay exact zeros
        __syncwarp();
        const int e0 = indptr[row], e1 = indptr[row + 1];
        for (int base = e0; base < e1; base += 32) {
            const int e = base + lane;
            const bool active = e < e1;
            const int c = active ? __ldg(code + indices[e]) : -1 - lane;  // inactive lanes never match
            const float v = active ? vals[e] : 0.f;
            unsigned g = __match_any_sync(kFull, c);
            const bool first = (g & ((1u << lane) - 1)) == 0;  // lowest lane of its group
            const int maxsz = __reduce_max_sync(kFull, __popc(g));
            float t = active ? buf[c] : 0.f;
            for (int k = 0; k < maxsz; ++k) {  // every member replays the group's adds in lane order
                const bool has = g != 0;
                const int src = has ? __ffs(g) - 1 : lane;
                g &= g - 1;
                const float vk = __shfl_sync(kFull, v, src);
                if (has) t += vk;
            }
            __syncwarp();
            if (active && first) buf[c] = t;
            __syncwarp();
        }
        if (lane == 0) buf[code[row + row_offset]] += diag[row];  // self term last (_nam.py:33)
        __syncwarp();
        float *o = out + row * ld;
        for (int c = lane; c < ld; c += 32) o[c] = buf[c];
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// general step, fp32 state, float4 lanes
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }

// QC = true: the batch-kurtosis QC statistic of the finished row (_nam.py:78-82: Pearson kurtosis
// across the per-batch means of s / C) is computed from the accumulators before the warp retires,
// which saves the separate pass over the state after the last diffusion step (<= 8 batches).
struct SpmmQc {
    const int8_t *col_batch;   // [ld] batch of each sample column, -1 for padding
    const double *inv_count;   // [ld] 1 / cells per sample (0 for padding)
    const double *batch_inv;   // [n_batches] 1 / samples per batch
    int n_batches;
    double *kurt;              // [n_rows]
};

template <int NV, int QC>  // QC = 0: plain step; QC = 2 / 4 / 8: batch-kurtosis epilogue for <= QC batches
__global__ void __launch_bounds__(256)
spmm_f32_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                const float *__restrict__ vals, const float *__restrict__ diag,
                const float *__restrict__ in, float *__restrict__ out, int64_t n_rows, int nvec,
                int64_t ld4, int64_t in_row_offset, SpmmQc qc) {
    int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const float4 *in4 = reinterpret_cast<const float4 *>(in);
    float4 *out4 = reinterpret_cast<float4 *>(out);
    int e0 = indptr[row], e1 = indptr[row + 1];
    float4 acc[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int base = e0; base < e1; base += 32) {
        int e = base + lane;
        int j = 0;
        float v = 0.f;
        if (e < e1) {
            j = indices[e];
            v = vals[e];
        }
        int cnt = min(32, e1 - base);
        for (int t = 0; t < cnt; t += 4) {
            int64_t jj[4];
            float vv[4];
            float4 x[4][NV];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                jj[u] = __shfl_sync(kFull, j, (t + u) & 31);
                vv[u] = __shfl_sync(kFull, v, (t + u) & 31);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    int c = lane + 32 * q;
                    x[u][q] = (t + u < cnt && c < nvec) ? ldg4(in4 + jj[u] * ld4 + c)
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (t + u < cnt) {
#pragma unroll
                    for (int q = 0; q < NV; ++q) {
                        acc[q].x = fmaf(vv[u], x[u][q].x, acc[q].x);
                        acc[q].y = fmaf(vv[u], x[u][q].y, acc[q].y);
                        acc[q].z = fmaf(vv[u], x[u][q].z, acc[q].z);
                        acc[q].w = fmaf(vv[u], x[u][q].w, acc[q].w);
                    }
                }
            }
        }
    }
    float d = diag[row];
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        int c = lane + 32 * q;
        if (c < nvec) {
            float4 x = ldg4(in4 + (row + in_row_offset) * ld4 + c);
            acc[q].x = fmaf(d, x.x, acc[q].x);
            acc[q].y = fmaf(d, x.y, acc[q].y);
            acc[q].z = fmaf(d, x.z, acc[q].z);
            acc[q].w = fmaf(d, x.w, acc[q].w);
            out4[row * ld4 + c] = acc[q];
        }
    }
    if (QC > 0) {
        constexpr int NB = QC > 0 ? QC : 1;
        double bs[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) bs[b] = 0.0;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            int c = lane + 32 * q;
            if (c < nvec) {
                const float v[4] = {acc[q].x, acc[q].y, acc[q].z, acc[q].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int col = 4 * c + k;
                    const int bid = qc.col_batch[col];
                    const double x = double(v[k]) * qc.inv_count[col];
#pragma unroll
                    for (int b = 0; b < NB; ++b) bs[b] += (bid == b) ? x : 0.0;
                }
            }
        }
        const int nb = qc.n_batches;
        double mm = 0.0;
#pragma unroll
        for (int b = 0; b < NB; ++b)
            if (b < nb) {
                bs[b] = warp_sum(bs[b]) * qc.batch_inv[b];  // mean over the batch's samples
                mm += bs[b];
            }
        if (lane == 0) {
            mm /= nb;
            double m2 = 0.0, m4 = 0.0;
#pragma unroll
            for (int b = 0; b < NB; ++b)
                if (b < nb) {
                    const double dlt = bs[b] - mm;
                    m2 += dlt * dlt;
                    m4 += dlt * dlt * dlt * dlt;
                }
            qc.kurt[row] = kurtosis_from_moments(mm, m2 / nb, m4 / nb, false);
        }
    }
}

// generic scalar kernel: any element type, any column count (public cna.tl.diffuse on user vectors)
template <typename T>
__global__ void spmm_generic_kernel(const int32_t *__restrict__ indptr,
                                    const int32_t *__restrict__ indices, const T *__restrict__ vals,
                                    const T *__restrict__ diag, const T *__restrict__ in,
                                    T *__restrict__ out, int64_t n_rows, int n_cols, int64_t ld,
                                    int64_t in_row_offset) {
    int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    int e0 = indptr[row], e1 = indptr[row + 1];
    T d = diag[row];
    for (int c = lane; c < n_cols; c += 32) {
        T acc = T(0);
        for (int e = e0; e < e1; ++e) acc += vals[e] * in[int64_t(indices[e]) * ld + c];
        out[row * ld + c] = acc + d * in[(row + in_row_offset) * ld + c];
    }
}

// ---------------------------------------------------------------------------------------------
// per-cell kurtosis across samples (auto-stop rule, _nam.py:59)
// ---------------------------------------------------------------------------------------------
__global__ void row_kurtosis_kernel(const float *__restrict__ s, int64_t ld, int64_t n_rows,
                                    int n_samples, const double *__restrict__ inv_count,
                                    double *__restrict__ kurt) {
    int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const float *p = s + row * ld;
    double sum = 0.0;
    for (int c = lane; c < n_samples; c += 32) sum += double(p[c]) * inv_count[c];
    double mean = warp_sum(sum) / n_samples;
    double s2 = 0.0, s4 = 0.0;
    for (int c = lane; c < n_samples; c += 32) {
        double dlt = double(p[c]) * inv_count[c] - mean;
        double d2 = dlt * dlt;
        s2 += d2;
        s4 += d2 * d2;
    }
    s2 = warp_sum(s2) / n_samples;
    s4 = warp_sum(s4) / n_samples;
    if (lane == 0) kurt[row] = kurtosis_from_moments(mean, s2, s4, true);
}

static inline unsigned warp_rows_grid(int64_t n_rows, int threads) {
    int64_t warps_per_block = threads / 32;
    return unsigned((n_rows + warps_per_block - 1) / warps_per_block);
}

}  // namespace cna

using namespace cna;

extern "C" {

int cna_graph_colsum(const int32_t *indptr, const int32_t *indices, const void *data, int is_f64,
                     int64_t n_rows, double *colsum, void *stream) {
    CNA_REQUIRE(n_rows >= 0 && indptr && colsum, "cna_graph_colsum: bad arguments");
    if (n_rows == 0) return CNA_OK;
    unsigned grid = warp_rows_grid(n_rows, 256);
    if (is_f64)
        colsum_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(
            indptr, indices, static_cast<const double *>(data), n_rows, colsum);
    else
        colsum_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(
            indptr, indices, static_cast<const float *>(data), n_rows, colsum);
    CNA_LAUNCHED("colsum_kernel");
    return CNA_OK;
}

int cna_graph_scale(const int32_t *indptr, const int32_t *indices, const void *data, int is_f64,
                    int64_t n_rows, const double *colsum, double self_weight, void *vals,
                    void *diag, int out_f64, int64_t row_offset, void *stream) {
    CNA_REQUIRE(n_rows >= 0 && indptr && colsum && vals && diag, "cna_graph_scale: bad arguments");
    if (n_rows == 0) return CNA_OK;
    unsigned grid = warp_rows_grid(n_rows, 256);
    cudaStream_t st = as_stream(stream);
#define CNA_SCALE(T, O)                                                                       \
    scale_kernel<T, O><<<grid, 256, 0, st>>>(indptr, indices, static_cast<const T *>(data),   \
                                              n_rows, colsum, self_weight,                    \
                                              static_cast<O *>(vals), static_cast<O *>(diag), row_offset)
    if (is_f64 && out_f64) CNA_SCALE(double, double);
    else if (is_f64) CNA_SCALE(double, float);
    else if (out_f64) CNA_SCALE(float, double);
    else CNA_SCALE(float, float);
#undef CNA_SCALE
    CNA_LAUNCHED("scale_kernel");
    return CNA_OK;
}

int cna_diffuse_onehot(const int32_t *indptr, const int32_t *indices, const float *vals,
                       const float *diag, const int32_t *code, int64_t n_rows, int n_samples,
                       float *out, int64_t ld, int64_t row_offset, void *stream) {
    CNA_REQUIRE(n_rows >= 0 && n_samples > 0 && ld >= n_samples, "cna_diffuse_onehot: bad shape");
    CNA_REQUIRE(ld <= 1024, "cna_diffuse_onehot: at most 1024 sample columns (got ld=%lld)",
                (long long)ld);
    if (n_rows == 0) return CNA_OK;
    cudaStream_t st = as_stream(stream);
    int64_t blocks = (n_rows + 7) / 8, cap = int64_t(num_sms()) * 8;
    unsigned grid = unsigned(blocks < cap ? blocks : cap);
    size_t smem = sizeof(float) * 8 * size_t(ld);
    onehot_step_kernel<<<grid, 256, smem, st>>>(indptr, indices, vals, diag, code, n_rows, out, ld, row_offset);
    CNA_LAUNCHED("onehot_step_kernel");
    return CNA_OK;
}

int cna_diffuse_step_f32(const int32_t *indptr, const int32_t *indices, const float *vals,
                         const float *diag, const float *in, float *out, int64_t n_rows,
                         int n_cols, int64_t ld, int64_t in_row_offset, void *stream) {
    CNA_REQUIRE(n_rows >= 0 && n_cols > 0 && ld >= n_cols, "cna_diffuse_step_f32: bad shape");
    CNA_REQUIRE(in != out, "cna_diffuse_step_f32: in-place step is not supported");
    if (n_rows == 0) return CNA_OK;
    unsigned grid = warp_rows_grid(n_rows, 256);
    cudaStream_t st = as_stream(stream);
    bool vec_ok = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) % 16 == 0);
    int nvec = (n_cols + 3) / 4;
    if (vec_ok && nvec <= 128) {
        int64_t ld4 = ld / 4;
        SpmmQc none{};
#define CNA_SPMM(NV) \
    spmm_f32_kernel<NV, 0><<<grid, 256, 0, st>>>(indptr, indices, vals, diag, in, out, n_rows, nvec, ld4, in_row_offset, none)
        if (nvec <= 32) CNA_SPMM(1);
        else if (nvec <= 64) CNA_SPMM(2);
        else if (nvec <= 96) CNA_SPMM(3);
        else CNA_SPMM(4);
#undef CNA_SPMM
        CNA_LAUNCHED("spmm_f32_kernel");
    } else {
        spmm_generic_kernel<float><<<grid, 256, 0, st>>>(indptr, indices, vals, diag, in, out,
                                                         n_rows, n_cols, ld, in_row_offset);
        CNA_LAUNCHED("spmm_generic_kernel<float>");
    }
    return CNA_OK;
}

int cna_diffuse_step_f32_qc(const int32_t *indptr, const int32_t *indices, const float *vals,
                            const float *diag, const float *in, float *out, int64_t n_rows, int n_cols,
                            int64_t ld, int64_t in_row_offset, const int8_t *col_batch,
                            const double *inv_count, const double *batch_inv, int n_batches, double *kurt,
                            void *stream) {
    CNA_REQUIRE(n_rows >= 0 && n_cols > 0 && ld >= n_cols && ld % 4 == 0, "cna_diffuse_step_f32_qc: bad shape");
    CNA_REQUIRE(in != out, "cna_diffuse_step_f32_qc: in-place step is not supported");
    CNA_REQUIRE(n_batches >= 2 && n_batches <= 8 && col_batch && inv_count && batch_inv && kurt,
                "cna_diffuse_step_f32_qc: 2..8 batches supported (got %d)", n_batches);
    CNA_REQUIRE((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) % 16 == 0,
                "cna_diffuse_step_f32_qc: state must be 16-byte aligned");
    int nvec = (n_cols + 3) / 4;
    CNA_REQUIRE(nvec <= 128, "cna_diffuse_step_f32_qc: at most 512 sample columns");
    if (n_rows == 0) return CNA_OK;
    unsigned grid = warp_rows_grid(n_rows, 256);
    cudaStream_t st = as_stream(stream);
    int64_t ld4 = ld / 4;
    SpmmQc qc{col_batch, inv_count, batch_inv, n_batches, kurt};
#define CNA_SPMM_NB(NV, NB) \
    spmm_f32_kernel<NV, NB><<<grid, 256, 0, st>>>(indptr, indices, vals, diag, in, out, n_rows, nvec, ld4, in_row_offset, qc)
#define CNA_SPMM(NV)                               \
    do {                                           \
        if (n_batches <= 2) CNA_SPMM_NB(NV, 2);    \
        else if (n_batches <= 4) CNA_SPMM_NB(NV, 4); \
        else CNA_SPMM_NB(NV, 8);                   \
    } while (0)
    if (nvec <= 32) CNA_SPMM(1);
    else if (nvec <= 64) CNA_SPMM(2);
    else if (nvec <= 96) CNA_SPMM(3);
    else CNA_SPMM(4);
#undef CNA_SPMM
#undef CNA_SPMM_NB
    CNA_LAUNCHED("spmm_f32_kernel<QC>");
    return CNA_OK;
}

int cna_diffuse_step_f64(const int32_t *indptr, const int32_t *indices, const double *vals,
                         const double *diag, const double *in, double *out, int64_t n_rows,
                         int n_cols, int64_t ld, int64_t in_row_offset, void *stream) {
    CNA_REQUIRE(n_rows >= 0 && n_cols > 0 && ld >= n_cols, "cna_diffuse_step_f64: bad shape");
    CNA_REQUIRE(in != out, "cna_diffuse_step_f64: in-place step is not supported");
    if (n_rows == 0) return CNA_OK;
    unsigned grid = warp_rows_grid(n_rows, 256);
    spmm_generic_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(indptr, indices, vals, diag,
                                                                     in, out, n_rows, n_cols, ld, in_row_offset);
    CNA_LAUNCHED("spmm_generic_kernel<double>");
    return CNA_OK;
}

int cna_row_kurtosis(const float *s, int64_t ld, int64_t n_rows, int n_samples,
                     const double *inv_count, double *kurt, void *stream) {
    CNA_REQUIRE(n_rows >= 0 && n_samples > 0 && ld >= n_samples, "cna_row_kurtosis: bad shape");
    if (n_rows == 0) return CNA_OK;
    unsigned grid = warp_rows_grid(n_rows, 256);
    row_kurtosis_kernel<<<grid, 256, 0, as_stream(stream)>>>(s, ld, n_rows, n_samples, inv_count, kurt);
    CNA_LAUNCHED("row_kurtosis_kernel");
    return CNA_OK;
}

}  // extern "C"
