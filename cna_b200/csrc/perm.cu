// Permutation engine: all Nnull shuffled-phenotype fits against the top NAM-PCs, one warp per
// permutation, fp64 throughout (1 - r^2 cancels to ~1e-4, so fp32 is not an option here).
//
// Reference: src/cna/tools/_association.py:35-61 (_reg / _stats / _minp_stats) called once per
// permutation from the Python loop at :84, and :94-97 (conditioned null phenotypes for the
// neighbourhood-level test).  The F survival function (scipy.special.fdtrc, :46) and the argmin over
// ks stay with the caller: they are O(Nnull * len(ks)) scalar work.
#include <cuda_fp16.h>

#include "common.cuh"

namespace cna {

template <int NQ>
__global__ void __launch_bounds__(256)
perm_stats_kernel(const double *__restrict__ y, const int32_t *__restrict__ perm, int64_t K, int n,
                  const double *__restrict__ C, const double *__restrict__ W, int r,
                  const double *__restrict__ Ut, int kmax, const int32_t *__restrict__ ks, int nks,
                  double *__restrict__ ssered, double *__restrict__ ssefull,
                  float *__restrict__ ycond, int64_t ld_y, int n_local, __half *__restrict__ yt_hi,
                  __half *__restrict__ yt_lo, int64_t ld16, int stage_u, int stage_wc) {
    // stage_u / stage_wc: the PC matrix / the design matrices fit in shared memory; otherwise they are
    // read where they lie (L2-resident: every warp of every CTA streams the same few hundred KB)
    extern __shared__ double sm[];
    double *ys = sm;               // [n]
    double *proj = ys + n;         // [warps][r]
    int *kss = reinterpret_cast<int *>(proj + (blockDim.x >> 5) * r);  // [nks], padded to 8 bytes
    double *next_free = reinterpret_cast<double *>(kss + ((nks + 1) & ~1));
    const double *Ws = W, *Cs = nullptr, *Us = Ut;
    for (int t = threadIdx.x; t < n; t += blockDim.x) ys[t] = y[t];
    if (stage_wc) {
        double *w_s = next_free, *c_s = w_s + r * n;  // [r][n] each (C transposed)
        next_free = c_s + r * n;
        for (int t = threadIdx.x; t < r * n; t += blockDim.x) {
            w_s[t] = W[t];
            c_s[t] = C[(t % n) * r + t / n];
        }
        Ws = w_s;
        Cs = c_s;
    }
    if (stage_u) {
        double *u_s = next_free;  // [kmax][n]
        for (int t = threadIdx.x; t < kmax * n; t += blockDim.x) u_s[t] = Ut[t];
        Us = u_s;
    }
    for (int t = threadIdx.x; t < nks; t += blockDim.x) kss[t] = ks[t];
    __syncthreads();

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, warps = blockDim.x >> 5;
    double *pw = proj + w * r;
    const double dn = double(n);
    for (int64_t k = int64_t(blockIdx.x) * warps + w; k < K; k += int64_t(gridDim.x) * warps) {
        double z[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            int m = lane + 32 * q;
            z[q] = (m < n) ? ys[perm[k * n + m]] : 0.0;  // _stats.py:18  Y[bix]
        }
        // zc = M z with M = I - C.W  (_association.py:51)
        for (int rr = 0; rr < r; ++rr) {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                int m = lane + 32 * q;
                if (m < n) acc += Ws[rr * n + m] * z[q];
            }
            acc = warp_sum(acc);
            if (lane == 0) pw[rr] = acc;
        }
        __syncwarp();
        for (int rr = 0; rr < r; ++rr) {
            double pr = pw[rr];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                int m = lane + 32 * q;
                if (m < n) z[q] -= (Cs ? Cs[rr * n + m] : C[m * r + rr]) * pr;
            }
        }
        __syncwarp();
        // zc /= zc.std(ddof=1)  (_association.py:52 — a pandas Series, hence ddof=1)
        double s1 = 0.0;
#pragma unroll
        for (int q = 0; q < NQ; ++q) s1 += z[q];
        double mean = warp_sum(s1) / dn;
        double s2 = 0.0;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            int m = lane + 32 * q;
            double d = (m < n) ? z[q] - mean : 0.0;
            s2 += d * d;
        }
        double sd = sqrt(warp_sum(s2) / (dn - 1.0));
        double sr = 0.0;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            z[q] = z[q] / sd;
            int m = lane + 32 * q;
            if (m < n) {
                sr += z[q] * z[q];
                if (k < n_local) {
                    if (ycond) ycond[int64_t(m) * ld_y + k] = float(z[q]);
                    if (yt_hi) {  // transposed fp16 hi/lo planes: the B operand of the tensor-core null GEMM
                        __half h = __float2half_rn(float(z[q]));
                        yt_hi[k * ld16 + m] = h;
                        yt_lo[k * ld16 + m] = __float2half_rn(float(z[q] - double(__half2float(h))));
                    }
                }
            } else {
                z[q] = 0.0;
            }
        }
        sr = warp_sum(sr);  // ssered = zc.zc  (_association.py:43)
        if (lane == 0 && ssered) ssered[k] = sr;
        // residual after regressing on the first j PCs, evaluated at j in ks (_association.py:35-42)
        int next = 0;
        for (int j = 0; j < kmax && next < nks; ++j) {
            double b = 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                int m = lane + 32 * q;
                if (m < n) b += Us[j * n + m] * z[q];
            }
            b = warp_sum(b);  // beta_j = U[:, j] . zc  (U orthonormal: independent of the residual)
            double ss = 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                int m = lane + 32 * q;
                if (m < n) {
                    z[q] -= b * Us[j * n + m];
                    ss += z[q] * z[q];
                }
            }
            if (j + 1 == kss[next]) {
                ss = warp_sum(ss);
                if (lane == 0) ssefull[k * nks + next] = ss;
                ++next;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// F survival function and min-p over ks for every permutation (_association.py:45-46, :53-60)
// ---------------------------------------------------------------------------------------------
// Continued fraction of the incomplete beta function (modified Lentz), evaluated in fp64.
__device__ double betacf(double a, double b, double x) {
    const double tiny = 1e-300, eps = 1e-16;
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < tiny) d = tiny;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 500; ++m) {
        const double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < eps) break;
    }
    return h;
}

// Regularised incomplete beta I_x(a, b).
__device__ double reg_inc_beta(double a, double b, double x) {
    if (!(x > 0.0)) return 0.0;
    if (!(x < 1.0)) return 1.0;
    const double lbt = lgamma(a + b) - lgamma(a) - lgamma(b) + a * log(x) + b * log1p(-x);
    const double bt = exp(lbt);
    if (x < (a + 1.0) / (a + b + 2.0)) return bt * betacf(a, b, x) / a;
    return 1.0 - bt * betacf(b, a, 1.0 - x) / b;
}

// scipy.special.fdtrc(dfn, dfd, f) = I_{dfd / (dfd + dfn f)}(dfd / 2, dfn / 2); 1 for f <= 0, NaN in -> NaN
__device__ double f_survival(double dfn, double dfd, double f) {
    if (f != f) return f;
    if (!(f > 0.0)) return 1.0;
    if (isinf(f)) return 0.0;
    return reg_inc_beta(0.5 * dfd, 0.5 * dfn, dfd / (dfd + dfn * f));
}

__global__ void minp_kernel(const double *__restrict__ ssered, const double *__restrict__ ssefull, int64_t K,
                            const int32_t *__restrict__ ks, int nks, int n, int r, double *__restrict__ minp,
                            int32_t *__restrict__ argk, double *__restrict__ r2) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= K) return;
    const double red = ssered[i];
    double best = nan("");
    int best_a = -1;
    for (int a = 0; a < nks; ++a) {
        const double k = double(ks[a]), full = ssefull[i * nks + a];
        const double f = ((red - full) / k) / (full / double(n));  // :45 (divides by n, not dof)
        const double p = f_survival(k, double(n - (1 + r + ks[a])), f);
        if (p == p && (best_a < 0 || p < best)) {  // nanargmin: first minimum, NaNs skipped
            best = p;
            best_a = a;
        }
    }
    minp[i] = best;
    argk[i] = best_a;
    r2[i] = best_a >= 0 ? 1.0 - ssefull[i * nks + best_a] / red : nan("");
}

}  // namespace cna

using namespace cna;

extern "C" int cna_perm_minp(const double *ssered, const double *ssefull, int64_t K, const int32_t *ks, int nks,
                             int n, int r, double *minp, int32_t *argk, double *r2, void *stream) {
    CNA_REQUIRE(K >= 0 && nks >= 1 && ssered && ssefull && ks && minp && argk && r2, "cna_perm_minp: bad arguments");
    if (K == 0) return CNA_OK;
    minp_kernel<<<unsigned((K + 127) / 128), 128, 0, as_stream(stream)>>>(ssered, ssefull, K, ks, nks, n, r, minp, argk, r2);
    CNA_LAUNCHED("minp_kernel");
    return CNA_OK;
}

extern "C" int cna_perm_stats(const double *y, const int32_t *perm, int64_t K, int n, const double *C,
                              const double *W, int r, const double *Ut, int kmax, const int32_t *ks,
                              int nks, double *ssered, double *ssefull, float *ycond, int64_t ld_y,
                              int n_local, void *yt_hi, void *yt_lo, int64_t ld16, void *stream) {
    CNA_REQUIRE(K >= 0 && n >= 2 && n <= 1024 && r >= 0 && kmax >= 0 && kmax <= n && nks >= 0,
                "cna_perm_stats: bad shape (K=%lld n=%d r=%d kmax=%d nks=%d)", (long long)K, n, r, kmax, nks);
    CNA_REQUIRE(kmax == 0 || (Ut && ks && nks >= 1 && ssefull), "cna_perm_stats: PCs requested but Ut/ks/ssefull missing");
    CNA_REQUIRE(n_local == 0 || ((ycond && ld_y >= n_local) || (yt_hi && yt_lo && ld16 >= n)),
                "cna_perm_stats: conditioned-phenotype buffer missing or too small");
    if (kmax == 0) nks = 0;
    if (K == 0) return CNA_OK;
    const int threads = 256, warps = threads / 32;
    // y, the per-warp projections and ks always live in shared memory; the design matrices (2 r n
    // doubles) and the PCs (kmax n doubles) only while they fit: beyond ~570 samples with the default
    // ks the PCs are streamed from L2 instead
    const size_t limit = 200 * 1024;
    size_t smem = sizeof(double) * (size_t(n) + size_t(warps) * r) + sizeof(int) * size_t((nks + 1) & ~1);
    const size_t wc_bytes = sizeof(double) * 2 * size_t(r) * n, u_bytes = sizeof(double) * size_t(kmax) * n;
    const int stage_wc = (r > 0 && smem + wc_bytes <= limit) ? 1 : 0;
    if (stage_wc) smem += wc_bytes;
    const int stage_u = (kmax > 0 && smem + u_bytes <= limit) ? 1 : 0;
    if (stage_u) smem += u_bytes;
    int64_t blocks = (K + warps - 1) / warps;
    int64_t cap = int64_t(num_sms()) * 2;
    unsigned grid = unsigned(blocks < cap ? blocks : cap);
    int nq = (n + 31) / 32;
    cudaStream_t st = as_stream(stream);
#define CNA_PERM(NQ)                                                                                  \
    do {                                                                                              \
        CNA_CUDA(cudaFuncSetAttribute(perm_stats_kernel<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      int(smem)));                                                    \
        perm_stats_kernel<NQ><<<grid, threads, smem, st>>>(y, perm, K, n, C, W, r, Ut, kmax, ks, nks, \
                                                            ssered, ssefull, ycond, ld_y, n_local,    \
                                                            static_cast<__half *>(yt_hi),             \
                                                            static_cast<__half *>(yt_lo), ld16,       \
                                                            stage_u, stage_wc);                       \
    } while (0)
    if (nq <= 2) CNA_PERM(2);
    else if (nq <= 4) CNA_PERM(4);
    else if (nq <= 8) CNA_PERM(8);
    else if (nq <= 16) CNA_PERM(16);
    else CNA_PERM(32);
#undef CNA_PERM
    CNA_LAUNCHED("perm_stats_kernel");
    return CNA_OK;
}
