// fp32 CUDA-core versions of the three dense contractions of the path.  They are the always-correct
// baseline (and the parity cross-check) for the tcgen05 kernels in gemm_tc.cu.
//
//   gram_simt_kernel        G += X^T X                   (_nam.py:105)
//   xb_simt_kernel<STORE>   out = X . B                  (_nam.py:106, V = NAM^T U / sqrt(svs))
//   xb_simt_kernel<HIST>    hist of (X . Ycond / n)^2    (_association.py:99 + _stats.py:52-54)
//
// All use 128 x 128 output tiles, 256 threads, an 8 x 8 micro-tile per thread split as 4+4 in both
// directions so that shared-memory reads are conflict-free float4s.
#include "common.cuh"

namespace cna {

constexpr int kTile = 128;
constexpr int kKT = 16;

__device__ __forceinline__ void fma_8x8(float (&acc)[8][8], const float4 &a0, const float4 &a1,
                                        const float4 &b0, const float4 &b1) {
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
}

// ---------------------------------------------------------------------------------------------
// Gram: contraction over cells.  blockIdx.x = row chunk, blockIdx.y = upper-triangular tile pair.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gram_simt_kernel(const float *__restrict__ x, int64_t ld, int64_t n_rows, int n, int rows_per_cta,
                 int n_tiles, double *__restrict__ gram) {
    __shared__ __align__(16) float As[kKT][kTile];
    __shared__ __align__(16) float Bs[kKT][kTile];
    int ta = 0, tb = 0;
    {  // decode the pair index: pairs are (0,0),(0,1),...,(0,T-1),(1,1),...
        int p = blockIdx.y;
        while (p >= n_tiles - ta) {
            p -= n_tiles - ta;
            ++ta;
        }
        tb = ta + p;
    }
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    int64_t row0 = int64_t(blockIdx.x) * rows_per_cta;
    int64_t row_end = row0 + rows_per_cta;
    if (row_end > n_rows) row_end = n_rows;
    // loader mapping: thread -> (row kk = tid / 32 + {0, 8}, float4 column c4 = tid % 32)
    const int lk = tid >> 5, lc = (tid & 31) * 4;
    for (int64_t k0 = row0; k0 < row_end; k0 += kKT) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int kk = lk + 8 * h;
            int64_t row = k0 + kk;
            float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
            if (row < row_end) {
                int ca = ta * kTile + lc, cb = tb * kTile + lc;
                if (ca < ld) va = __ldg(reinterpret_cast<const float4 *>(x + row * ld + ca));
                if (cb < ld) vb = __ldg(reinterpret_cast<const float4 *>(x + row * ld + cb));
            }
            *reinterpret_cast<float4 *>(&As[kk][lc]) = va;
            *reinterpret_cast<float4 *>(&Bs[kk][lc]) = vb;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kKT; ++kk) {
            float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4 *>(&As[kk][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4 *>(&Bs[kk][64 + tx * 4]);
            fma_8x8(acc, a0, a1, b0, b1);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int gi = ta * kTile + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
        if (gi >= n) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int gj = tb * kTile + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
            if (gj >= n) continue;
            double v = double(acc[i][j]);
            atomicAdd(gram + int64_t(gi) * n + gj, v);
            if (ta != tb) atomicAdd(gram + int64_t(gj) * n + gi, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// X . B with a store or a histogram epilogue.  A CTA owns 128 cells and walks all column tiles.
// ---------------------------------------------------------------------------------------------
enum class Epi { STORE, HIST };

struct XbArgs {
    const float *x;
    int64_t ld_x;
    int64_t n_rows;
    int n;            // contraction length (columns of X actually used; rest is zero padding)
    const float *b;   // [>= n rows][ld_b]
    int64_t ld_b;
    int n_out;        // number of output columns
    // STORE
    float *out;
    int64_t ld_out;
    // HIST
    const double *edges;
    int n_edges;
    uint32_t *hist;
    double inv_n;
    float reject_below;  // acc^2 below this can never reach edges[0]
};

template <Epi E>
__global__ void __launch_bounds__(256) xb_simt_kernel(XbArgs a) {
    __shared__ __align__(16) float As[kKT][kTile + 4];  // transposed X tile: As[k][cell]
    __shared__ __align__(16) float Bs[kKT][kTile];
    extern __shared__ double edges_s[];                 // HIST only
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    if (E == Epi::HIST) {
        for (int t = tid; t < a.n_edges; t += blockDim.x) edges_s[t] = a.edges[t];
        __syncthreads();
    }
    const int64_t row0 = int64_t(blockIdx.x) * kTile;
    const int k_end = (a.n + kKT - 1) / kKT * kKT;
    // A loader: thread -> cell ar = tid / 4 (+64), k offset ak = (tid % 4) * 4
    const int ar = tid >> 2, ak = (tid & 3) * 4;
    // B loader: thread -> k row bk = tid / 32 (+8), column bc = (tid % 32) * 4
    const int bk = tid >> 5, bc = (tid & 31) * 4;

    for (int c0 = 0; c0 < a.n_out; c0 += kTile) {
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        for (int k0 = 0; k0 < k_end; k0 += kKT) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int cell = ar + 64 * h;
                int64_t row = row0 + cell;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < a.n_rows && k0 + ak < a.ld_x)
                    v = __ldg(reinterpret_cast<const float4 *>(a.x + row * a.ld_x + k0 + ak));
                As[ak + 0][cell] = v.x;
                As[ak + 1][cell] = v.y;
                As[ak + 2][cell] = v.z;
                As[ak + 3][cell] = v.w;
                int kk = bk + 8 * h;
                float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k0 + kk < a.n && c0 + bc < a.ld_b)
                    w = __ldg(reinterpret_cast<const float4 *>(a.b + int64_t(k0 + kk) * a.ld_b + c0 + bc));
                *reinterpret_cast<float4 *>(&Bs[kk][bc]) = w;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < kKT; ++kk) {
                float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
                float4 a1 = *reinterpret_cast<const float4 *>(&As[kk][64 + ty * 4]);
                float4 b0 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
                float4 b1 = *reinterpret_cast<const float4 *>(&Bs[kk][64 + tx * 4]);
                fma_8x8(acc, a0, a1, b0, b1);
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int64_t row = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
            if (row >= a.n_rows) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int col = c0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
                if (col >= a.n_out) continue;
                float v = acc[i][j];
                if (E == Epi::STORE) {
                    a.out[row * a.ld_out + col] = v;
                } else {
                    if (v * v < a.reject_below) continue;
                    double z = double(v) * a.inv_n;
                    double z2 = z * z;
                    if (!(z2 >= edges_s[0])) continue;
                    int lo = 0, hi = a.n_edges - 1;  // largest b with edges[b] <= z2
                    while (lo < hi) {
                        int mid = (lo + hi + 1) >> 1;
                        if (edges_s[mid] <= z2) lo = mid;
                        else hi = mid - 1;
                    }
                    atomicAdd(a.hist + int64_t(col) * a.n_edges + lo, 1u);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// observed-coefficient histograms, max |ncorr|, per-cell fdr lookup
// ---------------------------------------------------------------------------------------------
__global__ void absmax_kernel(const double *__restrict__ v, const uint8_t *__restrict__ valid,
                              int64_t n, double *out) {
    double m = 0.0;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += int64_t(gridDim.x) * blockDim.x)
        if (!valid || valid[i]) m = fmax(m, fabs(v[i]));
    m = warp_max(m);
    // non-negative doubles order like their bit patterns
    if ((threadIdx.x & 31) == 0 && m > 0.0)
        atomicMax(reinterpret_cast<unsigned long long *>(out), (unsigned long long)__double_as_longlong(m));
}

__device__ __forceinline__ int upper_le(const double *t, int n, double v) {
    // number of entries <= v in ascending t (np.searchsorted(t, v, side='right'))
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (t[mid] <= v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__global__ void obs_hist_kernel(const double *__restrict__ ncorr, const uint8_t *__restrict__ valid,
                                int64_t n, const double *__restrict__ edges,
                                const double *__restrict__ thr, int n_edges_cap,
                                const int32_t *__restrict__ n_edges_dev, uint32_t *rank_hist,
                                uint32_t *det_hist) {
    // n_edges_dev != NULL: the number of edges is whatever cna_fdr_thresholds left on the device
    // (n_edges_cap then only sizes the shared-memory tables)
    extern __shared__ double sm[];
    double *e = sm, *t = sm + n_edges_cap;
    uint32_t *hr = reinterpret_cast<uint32_t *>(t + n_edges_cap), *hd = hr + n_edges_cap;
    const int n_edges = n_edges_dev ? min(__ldg(n_edges_dev), n_edges_cap) : n_edges_cap;
    for (int i = threadIdx.x; i < n_edges; i += blockDim.x) {
        e[i] = edges[i];
        t[i] = thr[i];
        hr[i] = 0;
        hd[i] = 0;
    }
    __syncthreads();
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += int64_t(gridDim.x) * blockDim.x) {
        if (valid && !valid[i]) continue;
        double c = ncorr[i];
        int br = upper_le(e, n_edges, c * c);  // edges[br-1] <= c^2
        if (br > 0) atomicAdd(hr + br - 1, 1u);
        // strict: thresholds[b] < |c|  <=>  b < #{thresholds < |c|}
        double ac = fabs(c);
        int lo = 0, hi = n_edges;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (t[mid] < ac) lo = mid + 1;
            else hi = mid;
        }
        if (lo > 0) atomicAdd(hd + lo - 1, 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_edges; i += blockDim.x) {
        if (hr[i]) atomicAdd(rank_hist + i, hr[i]);
        if (hd[i]) atomicAdd(det_hist + i, hd[i]);
    }
}

__global__ void cell_fdr_kernel(const double *__restrict__ ncorr, const uint8_t *__restrict__ valid,
                                int64_t n, const double *__restrict__ thr,
                                const double *__restrict__ pmin, int n_thr_cap, const int32_t *__restrict__ n_thr_dev,
                                double *coef, double *fdr) {
    extern __shared__ double sm[];
    double *t = sm, *p = sm + n_thr_cap;
    const int n_thr = n_thr_dev ? min(__ldg(n_thr_dev), n_thr_cap) : n_thr_cap;
    for (int i = threadIdx.x; i < n_thr; i += blockDim.x) {
        t[i] = thr[i];
        p[i] = pmin[i];
    }
    __syncthreads();
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += int64_t(gridDim.x) * blockDim.x) {
        if (valid && !valid[i]) {
            coef[i] = nan("");
            fdr[i] = 1.0;
            continue;
        }
        double c = ncorr[i];
        coef[i] = c;
        int idx = upper_le(t, n_thr, fabs(c));
        fdr[i] = idx > 0 ? p[idx - 1] : 1.0;
    }
}

}  // namespace cna

using namespace cna;

extern "C" {

int cna_gram_simt(const float *x, int64_t ld_x, int64_t n_rows, int n, double *gram, void *stream) {
    CNA_REQUIRE(n_rows >= 0 && n > 0 && ld_x >= n && ld_x % 4 == 0, "cna_gram: bad shape (n=%d ld=%lld)", n, (long long)ld_x);
    CNA_REQUIRE(reinterpret_cast<uintptr_t>(x) % 16 == 0, "cna_gram: x must be 16-byte aligned");
    if (n_rows == 0) return CNA_OK;
    const int rows_per_cta = 2048;
    int n_tiles = (n + kTile - 1) / kTile;
    dim3 grid(unsigned((n_rows + rows_per_cta - 1) / rows_per_cta), unsigned(n_tiles * (n_tiles + 1) / 2));
    gram_simt_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, ld_x, n_rows, n, rows_per_cta, n_tiles, gram);
    CNA_LAUNCHED("gram_simt_kernel");
    return CNA_OK;
}

int cna_gram(const float *x, int64_t ld_x, int64_t n_rows, int n, double *gram, void *stream) {
    return cna_gram_simt(x, ld_x, n_rows, n, gram, stream);
}

int cna_right_multiply(const float *x, int64_t ld_x, int64_t n_rows, int n, const float *b,
                       int64_t ld_b, int n_out, float *out, int64_t ld_out, void *stream) {
    CNA_REQUIRE(n_rows >= 0 && n > 0 && ld_x >= n && ld_x % 4 == 0 && ld_b % 4 == 0 && n_out > 0 && ld_out >= n_out,
                "cna_right_multiply: bad shape");
    CNA_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(b)) % 16 == 0,
                "cna_right_multiply: operands must be 16-byte aligned");
    if (n_rows == 0) return CNA_OK;
    XbArgs a{};
    a.x = x; a.ld_x = ld_x; a.n_rows = n_rows; a.n = n; a.b = b; a.ld_b = ld_b; a.n_out = n_out;
    a.out = out; a.ld_out = ld_out;
    unsigned grid = unsigned((n_rows + kTile - 1) / kTile);
    xb_simt_kernel<Epi::STORE><<<grid, 256, 0, as_stream(stream)>>>(a);
    CNA_LAUNCHED("xb_simt_kernel<STORE>");
    return CNA_OK;
}

int cna_null_hist(const float *x, int64_t ld_x, int64_t n_rows, int n, const float *ycond,
                  int64_t ld_y, int n_null, const double *edges, int n_edges, double edge0,
                  uint32_t *hist, void *stream) {
    CNA_REQUIRE(n_rows >= 0 && n > 0 && ld_x >= n && ld_x % 4 == 0 && ld_y % 4 == 0 && ld_y >= n_null,
                "cna_null_hist: bad shape");
    CNA_REQUIRE(n_edges > 0 && n_edges <= 4096, "cna_null_hist: 1..4096 edges supported (got %d)", n_edges);
    CNA_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(ycond)) % 16 == 0,
                "cna_null_hist: operands must be 16-byte aligned");
    if (n_rows == 0 || n_null == 0) return CNA_OK;
    XbArgs a{};
    a.x = x; a.ld_x = ld_x; a.n_rows = n_rows; a.n = n; a.b = ycond; a.ld_b = ld_y; a.n_out = n_null;
    a.edges = edges; a.n_edges = n_edges; a.hist = hist; a.inv_n = 1.0 / double(n);
    // acc^2 < edge0 * n^2 (with a relative safety margin far above fp32 rounding) can never count
    double rb = edge0 * double(n) * double(n) * (1.0 - 1e-5);
    a.reject_below = rb > 0.0 ? float(rb) * (1.0f - 1e-6f) : 0.f;
    unsigned grid = unsigned((n_rows + kTile - 1) / kTile);
    size_t smem = sizeof(double) * n_edges;
    xb_simt_kernel<Epi::HIST><<<grid, 256, smem, as_stream(stream)>>>(a);
    CNA_LAUNCHED("xb_simt_kernel<HIST>");
    return CNA_OK;
}

int cna_absmax(const double *v, const uint8_t *row_valid, int64_t n_rows, double *out, void *stream) {
    if (n_rows <= 0) return CNA_OK;
    int64_t blocks = (n_rows + 255) / 256;
    unsigned grid = unsigned(blocks < 1184 ? blocks : 1184);
    absmax_kernel<<<grid, 256, 0, as_stream(stream)>>>(v, row_valid, n_rows, out);
    CNA_LAUNCHED("absmax_kernel");
    return CNA_OK;
}

int cna_obs_hist_dev(const double *ncorr, const uint8_t *row_valid, int64_t n_rows, const double *edges,
                     const double *thresholds, int n_edges, const int32_t *n_edges_dev, uint32_t *rank_hist,
                     uint32_t *det_hist, void *stream) {
    CNA_REQUIRE(n_edges > 0 && n_edges <= 2048, "cna_obs_hist: 1..2048 edges supported");
    if (n_rows <= 0) return CNA_OK;
    int64_t blocks = (n_rows + 255) / 256;
    unsigned grid = unsigned(blocks < 592 ? blocks : 592);
    size_t smem = (2 * sizeof(double) + 2 * sizeof(uint32_t)) * n_edges;
    obs_hist_kernel<<<grid, 256, smem, as_stream(stream)>>>(ncorr, row_valid, n_rows, edges, thresholds,
                                                          n_edges, n_edges_dev, rank_hist, det_hist);
    CNA_LAUNCHED("obs_hist_kernel");
    return CNA_OK;
}

int cna_obs_hist(const double *ncorr, const uint8_t *row_valid, int64_t n_rows, const double *edges,
                 const double *thresholds, int n_edges, uint32_t *rank_hist, uint32_t *det_hist,
                 void *stream) {
    return cna_obs_hist_dev(ncorr, row_valid, n_rows, edges, thresholds, n_edges, nullptr, rank_hist, det_hist,
                            stream);
}

int cna_cell_fdr_dev(const double *ncorr, const uint8_t *row_valid, int64_t n_rows,
                     const double *thresholds, const double *prefix_min_fdr, int n_thr, const int32_t *n_thr_dev,
                     double *coef, double *fdr, void *stream) {
    CNA_REQUIRE(n_thr > 0 && n_thr <= 2048, "cna_cell_fdr: 1..2048 thresholds supported");
    if (n_rows <= 0) return CNA_OK;
    int64_t blocks = (n_rows + 255) / 256;
    unsigned grid = unsigned(blocks < 1184 ? blocks : 1184);
    size_t smem = 2 * sizeof(double) * n_thr;
    cell_fdr_kernel<<<grid, 256, smem, as_stream(stream)>>>(ncorr, row_valid, n_rows, thresholds,
                                                          prefix_min_fdr, n_thr, n_thr_dev, coef, fdr);
    CNA_LAUNCHED("cell_fdr_kernel");
    return CNA_OK;
}

int cna_cell_fdr(const double *ncorr, const uint8_t *row_valid, int64_t n_rows,
                 const double *thresholds, const double *prefix_min_fdr, int n_thr, double *coef,
                 double *fdr, void *stream) {
    return cna_cell_fdr_dev(ncorr, row_valid, n_rows, thresholds, prefix_min_fdr, n_thr, nullptr, coef, fdr, stream);
}

}  // extern "C"
