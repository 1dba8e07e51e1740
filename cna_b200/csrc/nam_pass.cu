// Kernel (ii): everything between the diffusion and the Gram matrix that is local to one cell.
//
// In the cells x samples layout each neighbourhood is one row, so QC (batch kurtosis), the
// sample reindex/filter, centring, the ridge residualisation, the ddof=1 standardisation and the
// neighbourhood coefficient are all row-local: one read of the fp32 state, one write of the
// fp32 residualised NAM.  Arithmetic inside a row is fp64 (the reference is fp64 throughout and
// the per-row work is tiny: ~2 n (r+3) flops).
//
// Reference: src/cna/tools/_nam.py:78-99 (QC), :118-159 (_resid_nam), _association.py:175-185
// (reindex, filter, zero-variance drop), :77 (ncorrs).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

namespace cna {

// Pearson kurtosis across the per-batch means of the values in `rowbuf` (per-warp shared memory).
// Batch b owns positions seg_order[seg_off[b] .. seg_off[b+1]).  `means` is per-warp scratch [nb].
// Every lane returns the result.  (_nam.py:78-82)
__device__ __forceinline__ double batch_kurtosis_of_row(const double *rowbuf, double *means,
                                                        const int *seg_order, const int *seg_off,
                                                        int nb, int lane) {
    int G = 1;  // lanes cooperating on one batch
    while (G * 2 * nb <= 32) G *= 2;
    int groups = 32 / G, g = lane / G, u = lane % G;
    int iters = (nb + groups - 1) / groups;
    for (int it = 0; it < iters; ++it) {
        int b = g + it * groups;
        bool active = b < nb;
        double acc = 0.0;
        int t0 = 0, t1 = 0;
        if (active) {
            t0 = seg_off[b];
            t1 = seg_off[b + 1];
            for (int t = t0 + u; t < t1; t += G) acc += rowbuf[seg_order[t]];
        }
        for (int o = G >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
        if (active && u == 0) means[b] = acc / double(t1 - t0);
    }
    __syncwarp();
    double tot = 0.0;
    for (int b = lane; b < nb; b += 32) tot += means[b];
    double mm = warp_sum(tot) / nb;
    double s2 = 0.0, s4 = 0.0;
    for (int b = lane; b < nb; b += 32) {
        double d = means[b] - mm;
        double d2 = d * d;
        s2 += d2;
        s4 += d2 * d2;
    }
    s2 = warp_sum(s2) / nb;
    s4 = warp_sum(s4) / nb;
    __syncwarp();
    return kurtosis_from_moments(mm, s2, s4, false);
}

// ---------------------------------------------------------------------------------------------
// QC: batch kurtosis of the raw NAM over all samples
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
batch_kurtosis_kernel(const float *__restrict__ s, int64_t ld, int64_t n_rows,
                      const double *__restrict__ inv_count, const int32_t *__restrict__ seg_order_g,
                      const int32_t *__restrict__ seg_off_g, int nb, int n_sel,
                      double *__restrict__ kurt) {
    extern __shared__ double sm[];
    const int warps = blockDim.x >> 5;
    double *invc = sm;                       // [n_sel] scaling of the selected columns
    double *rowbuf = invc + n_sel;           // [warps][n_sel]
    double *means = rowbuf + warps * n_sel;  // [warps][nb]
    int *seg_col = reinterpret_cast<int *>(means + warps * nb);  // [n_sel] state column ids
    int *seg_pos = seg_col + n_sel;          // [n_sel] identity positions
    int *seg_off = seg_pos + n_sel;          // [nb + 1]
    for (int t = threadIdx.x; t < n_sel; t += blockDim.x) {
        int c = seg_order_g[t];
        seg_col[t] = c;
        seg_pos[t] = t;
        invc[t] = inv_count[c];
    }
    for (int t = threadIdx.x; t <= nb; t += blockDim.x) seg_off[t] = seg_off_g[t];
    __syncthreads();
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double *rb = rowbuf + w * n_sel, *mb = means + w * nb;
    int64_t stride = int64_t(gridDim.x) * warps;
    for (int64_t row = int64_t(blockIdx.x) * warps + w; row < n_rows; row += stride) {
        const float *p = s + row * ld;
        for (int t = lane; t < n_sel; t += 32) rb[t] = double(__ldg(p + seg_col[t])) * invc[t];
        __syncwarp();
        double k = batch_kurtosis_of_row(rb, mb, seg_pos, seg_off, nb, lane);
        if (lane == 0) kurt[row] = k;
    }
}

// ---------------------------------------------------------------------------------------------
// fused select / centre / residualise / standardise / ncorr pass: generic kernel (large n.(r + batches))
// ---------------------------------------------------------------------------------------------
// A warp owns R consecutive rows at a time: every W / C coefficient read from shared memory is used
// for R rows, and the R independent shuffle reductions overlap each other's latency.
// EXACT: NQ == ceil(n / 32), so only the last column group can run past n (the bounds checks of
// the others fold away at compile time).  RT > 0: r <= RT and the loops over the r design columns
// are fully unrolled (uniform `rr < r` guards); RT == 0: runtime loops.
#define CNA_IN(q, m) ((EXACT && (q) < NQ - 1) || (m) < n)
template <int NQ, int R, bool EXACT, int RT>
__global__ void __launch_bounds__(256) resid_kernel(cna_resid_args a) {
    extern __shared__ double sm[];
    const int warps = blockDim.x >> 5;
    const int n = a.n, r = a.r, nb = a.n_batches;
    const bool want_kurt = (a.kurt != nullptr) && nb > 1;
    double *Wt = sm;                         // [r][n]
    double *Ct = Wt + r * n;                 // [r][n]  (transposed copy of C [n x r])
    double *ys = Ct + r * n;                 // [n]
    double *invc = ys + n;                   // [n]
    double *proj = invc + n;                 // [warps][R][r]
    double *rowbuf = proj + warps * R * r;   // [warps][n]   (only when want_kurt)
    double *means = rowbuf + (want_kurt ? warps * n : 0);  // [warps][nb]
    int *colmap = reinterpret_cast<int *>(means + (want_kurt ? warps * nb : 0));  // [n]
    int *seg_order = colmap + n;             // [n]
    int *seg_off = seg_order + n;            // [nb + 1]
    for (int t = threadIdx.x; t < r * n; t += blockDim.x) {
        Wt[t] = a.Wt[t];
        int rr = t / n, m = t % n;
        Ct[t] = a.C[m * r + rr];
    }
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        int c = a.colmap[t];
        colmap[t] = c;
        ys[t] = a.y[t];
        invc[t] = a.inv_count[c];
        if (want_kurt) seg_order[t] = a.seg_order[t];
    }
    if (want_kurt)
        for (int t = threadIdx.x; t <= nb; t += blockDim.x) seg_off[t] = a.seg_off[t];
    __syncthreads();

    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double qc_thr = 0.0;
    if (a.qc_kurt) {  // _nam.py:94: max(6, 2 * median) with Python's max (a NaN median gives 6)
        const double two_med = 2.0 * a.qc_median[0];
        qc_thr = two_med > 6.0 ? two_med : 6.0;
    }
    double *pw = proj + w * R * r;
    double *rb = rowbuf + w * n, *mb = means + w * nb;
    const double dn = double(n);
    const int64_t stride = int64_t(gridDim.x) * warps * R;
    for (int64_t row0 = (int64_t(blockIdx.x) * warps + w) * R; row0 < a.n_rows; row0 += stride) {
        double x[R][NQ];
        bool valid[R];
        double sum[R], ss[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int64_t row = row0 + i;
            const float *p = a.s + (row < a.n_rows ? row : row0) * a.ld_s;
            sum[i] = 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                int m = lane + 32 * q;
                x[i][q] = CNA_IN(q, m) ? double(__ldg(p + colmap[m])) * invc[m] : 0.0;
                sum[i] += x[i][q];
            }
        }
#pragma unroll
        for (int i = 0; i < R; ++i) sum[i] = warp_sum(sum[i]) / dn;
#pragma unroll
        for (int i = 0; i < R; ++i) {
            ss[i] = 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                int m = lane + 32 * q;
                x[i][q] = CNA_IN(q, m) ? x[i][q] - sum[i] : 0.0;  // _nam.py:122
                ss[i] += x[i][q] * x[i][q];
            }
        }
#pragma unroll
        for (int i = 0; i < R; ++i) {
            double var0 = warp_sum(ss[i]) / (dn - 1.0);
            const int64_t row = row0 + i;
            bool keep = row < a.n_rows;
            if (keep && a.row_keep) keep = a.row_keep[row] != 0;
            else if (keep && a.qc_kurt) keep = a.qc_kurt[row] < qc_thr;  // NaN -> dropped, _nam.py:96
            valid[i] = keep && !(var0 == 0.0);  // _association.py:182-185
        }
        // rank-r update X <- X - (X Wt^T) C^T   (_nam.py:133-135 / :146-148 with M = I - C.W)
        auto project = [&](int rr) {
            double acc[R];
#pragma unroll
            for (int i = 0; i < R; ++i) acc[i] = 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                int m = lane + 32 * q;
                double wv = CNA_IN(q, m) ? Wt[rr * n + m] : 0.0;
#pragma unroll
                for (int i = 0; i < R; ++i) acc[i] += x[i][q] * wv;
            }
#pragma unroll
            for (int i = 0; i < R; ++i) acc[i] = warp_sum(acc[i]);
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < R; ++i) pw[i * r + rr] = acc[i];
            }
        };
        auto update = [&](int rr) {
            double pr[R];
#pragma unroll
            for (int i = 0; i < R; ++i) pr[i] = pw[i * r + rr];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                int m = lane + 32 * q;
                double cv = CNA_IN(q, m) ? Ct[rr * n + m] : 0.0;
#pragma unroll
                for (int i = 0; i < R; ++i) x[i][q] -= pr[i] * cv;
            }
        };
        if (RT > 0) {
#pragma unroll
            for (int rr = 0; rr < RT; ++rr)
                if (rr < r) project(rr);
        } else {
            for (int rr = 0; rr < r; ++rr) project(rr);
        }
        __syncwarp();
        if (RT > 0) {
#pragma unroll
            for (int rr = 0; rr < RT; ++rr)
                if (rr < r) update(rr);
        } else {
            for (int rr = 0; rr < r; ++rr) update(rr);
        }
        __syncwarp();
        if (want_kurt) {  // _nam.py:150-155
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const int64_t row = row0 + i;
                if (row >= a.n_rows) break;  // uniform over the warp
                if (!valid[i]) {
                    if (lane == 0) a.kurt[row] = nan("");
                    continue;
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    int m = lane + 32 * q;
                    if (CNA_IN(q, m)) rb[m] = x[i][q];
                }
                __syncwarp();
                double k = batch_kurtosis_of_row(rb, mb, seg_order, seg_off, nb, lane);
                if (lane == 0) a.kurt[row] = k;
                __syncwarp();
            }
        } else if (a.kurt && lane == 0) {
#pragma unroll
            for (int i = 0; i < R; ++i)
                if (row0 + i < a.n_rows) a.kurt[row0 + i] = nan("");
        }
        // ddof=1 standardisation (_nam.py:159; pandas std recomputes the mean)
        double s1[R], s2[R], dot[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            s1[i] = 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) s1[i] += x[i][q];
        }
#pragma unroll
        for (int i = 0; i < R; ++i) s1[i] = warp_sum(s1[i]) / dn;
#pragma unroll
        for (int i = 0; i < R; ++i) {
            s2[i] = 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                int m = lane + 32 * q;
                double d = CNA_IN(q, m) ? x[i][q] - s1[i] : 0.0;
                s2[i] += d * d;
            }
        }
#pragma unroll
        for (int i = 0; i < R; ++i) s2[i] = 1.0 / sqrt(warp_sum(s2[i]) / (dn - 1.0));  // 1 / std
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int64_t row = row0 + i;
            dot[i] = 0.0;
            if (row >= a.n_rows) continue;  // uniform over the warp
            float *o = a.x_out ? a.x_out + row * a.ld_x : nullptr;
            __half *ph = a.x16_hi ? static_cast<__half *>(a.x16_hi) + row * a.ld16 : nullptr;
            __half *pl = a.x16_hi ? static_cast<__half *>(a.x16_lo) + row * a.ld16 : nullptr;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                int m = lane + 32 * q;
                double v = (CNA_IN(q, m) && valid[i]) ? x[i][q] * s2[i] : 0.0;  // rows of dropped cells are zero
                if (CNA_IN(q, m)) dot[i] += v * ys[m];
                if (o && m < a.ld_x) o[m] = float(v);
                if (ph && m < a.ld16) {  // v = hi + lo to 2^-22
                    __half h = __float2half_rn(float(v));
                    ph[m] = h;
                    pl[m] = __float2half_rn(float(v - double(__half2float(h))));
                }
            }
        }
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int64_t row = row0 + i;
            if (row >= a.n_rows) continue;
            double d = warp_sum(dot[i]);
            if (lane == 0) {
                a.ncorr[row] = valid[i] ? d / dn : 0.0;  // _association.py:77
                a.row_valid[row] = valid[i] ? 1 : 0;
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// fused pass, formulation as linear functionals of the raw row (the default kernel)
// ---------------------------------------------------------------------------------------------
// Everything the pass needs from a row x (selected samples, scaled by 1/C) is a handful of linear
// functionals of x plus two sums of squares:
//     mean = 1.x / n            p = W.x  (r values; W.1 = 0 because the columns of C are centred)
//     bmean_b = (1/|b|) sum_{k in b} x_k        dot = y.x
//     x' = x - mean - C.p       (_nam.py:122, :133-148)        ss0 = |x - mean|^2,  ss = |x'|^2, s1 = 1.x'
// so the row is read once into registers (lanes own columns), the m1 = 2 + r + nb functionals are
// accumulated against one shared-memory table F [m1 x n] and reduced 16 at a time with a transposing
// butterfly (3 instructions per reduced value instead of 15 for one shuffle tree each), and the batch
// means / coefficient follow from the functionals:
//     batch mean of x' = bmean_b - mean - Cbar_b.p,     x'.y = dot - mean * sum(y) - (C^T y).p
// A warp owns R rows at a time, so every table entry read from shared memory serves R rows.
struct ResidTables {
    int m1;      // number of functionals: 0 = sum, 1 = y, 2 .. 2+r = W rows, then the batch means
    int nbk;     // number of batch rows (0 when the kurtosis is not wanted)
};

// Reduce v[0..16) across the warp: afterwards lane L holds the warp total of v[L >> 1] in v[0].
template <int HALF>
__device__ __forceinline__ void butterfly_step(double (&v)[16], int lane) {
    constexpr int BIT = 2 * HALF;
    const bool upper = (lane & BIT) != 0;
#pragma unroll
    for (int j = 0; j < HALF; ++j) {
        const double send = upper ? v[j] : v[j + HALF];
        const double keep = upper ? v[j + HALF] : v[j];
        v[j] = keep + __shfl_xor_sync(kFull, send, BIT);
    }
}
__device__ __forceinline__ void butterfly16(double (&v)[16], int lane) {
    butterfly_step<8>(v, lane);
    butterfly_step<4>(v, lane);
    butterfly_step<2>(v, lane);
    butterfly_step<1>(v, lane);
    v[0] += __shfl_xor_sync(kFull, v[0], 1);
}

template <int NQ, int R>
__global__ void __launch_bounds__(256, 2) resid_lin_kernel(cna_resid_args a, ResidTables tb) {
    // Every table row is padded to LDN = 32 NQ columns with zeros and the number of functionals to a
    // multiple of FPC with zero rows, and the padding columns of a row are read as 0 * s[row][colmap[0]],
    // so the inner loops carry no bounds checks and address the tables with compile-time offsets.
    constexpr int LDN = 32 * NQ;
    constexpr int FPC = 16 / R;  // functionals per butterfly
    extern __shared__ double sm[];
    const int warps = blockDim.x >> 5;
    const int n = a.n, r = a.r, m1 = tb.m1, nbk = tb.nbk;
    const int m1p = (m1 + FPC - 1) / FPC * FPC, tws = m1p + 3;
    double *F = sm;                          // [m1p][LDN]: ones, y, W rows, batch-mean rows
    double *Ct = F + m1p * LDN;              // [r][LDN]
    double *invc = Ct + r * LDN;             // [LDN]
    double *cbar = invc + LDN;               // [nbk][r] batch means of the columns of C
    double *cy = cbar + nbk * r;             // [r] C^T y, then [1] sum(y)
    double *tot = cy + r + 1;                // [warps][R][tws]
    int *colmap = reinterpret_cast<int *>(tot + warps * R * tws);  // [LDN]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;

    // ---- tables (a few thousand flops per CTA) ----
    for (int t = threadIdx.x; t < (m1p + r) * LDN; t += blockDim.x) F[t] = 0.0;  // F and Ct
    for (int t = threadIdx.x; t < LDN; t += blockDim.x) {
        const int c = t < n ? a.colmap[t] : a.colmap[0];
        colmap[t] = c;
        invc[t] = t < n ? a.inv_count[c] : 0.0;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        F[t] = 1.0;
        F[LDN + t] = a.y[t];
    }
    for (int t = threadIdx.x; t < r * n; t += blockDim.x) {
        const int rr = t / n, m = t % n;
        F[(2 + rr) * LDN + m] = a.Wt[t];
        Ct[rr * LDN + m] = a.C[m * r + rr];
    }
    for (int b = w; b < nbk; b += warps) {
        const int t0 = a.seg_off[b], t1 = a.seg_off[b + 1];
        const double inv = 1.0 / double(t1 - t0);
        for (int t = t0 + lane; t < t1; t += 32) F[(2 + r + b) * LDN + a.seg_order[t]] = inv;
        for (int rr = 0; rr < r; ++rr) {
            double acc = 0.0;
            for (int t = t0 + lane; t < t1; t += 32) acc += a.C[a.seg_order[t] * r + rr];
            acc = warp_sum(acc);
            if (lane == 0) cbar[b * r + rr] = acc * inv;
        }
    }
    for (int rr = w; rr <= r; rr += warps) {  // rr == r: sum(y)
        double acc = 0.0;
        for (int t = lane; t < n; t += 32) acc += (rr < r ? a.C[t * r + rr] : 1.0) * a.y[t];
        acc = warp_sum(acc);
        if (lane == 0) cy[rr] = acc;
    }
    __syncthreads();

    double *tw = tot + w * R * tws;
    const double dn = double(n), inv_n = 1.0 / dn, inv_nm1 = 1.0 / (dn - 1.0);
    double thr = 0.0;
    if (a.qc_kurt) {  // _nam.py:94: threshold = max(6, 2 * median) with Python's max (NaN -> 6)
        const double two_med = 2.0 * a.qc_median[0];
        thr = two_med > 6.0 ? two_med : 6.0;
    }
    const int *cml = colmap + lane;
    const double *icl = invc + lane;
    const double *Fl = F + lane, *Cl = Ct + lane;
    const int stride = gridDim.x * warps * R;
    for (int64_t row0 = int64_t(blockIdx.x * warps + w) * R; row0 < a.n_rows; row0 += stride) {
        double x[R][NQ];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int64_t row = row0 + i < a.n_rows ? row0 + i : row0;
            const float *p = a.s + row * a.ld_s;
#pragma unroll
            for (int q = 0; q < NQ; ++q) x[i][q] = double(__ldg(p + cml[32 * q])) * icl[32 * q];
        }
        // ---- the functionals, 16 (= FPC per row) at a time ----
        for (int f0 = 0; f0 < m1p; f0 += FPC) {
            double acc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0.0;
            const double *Fp = Fl + f0 * LDN;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
#pragma unroll
                for (int f = 0; f < FPC; ++f) {
                    const double c = Fp[f * LDN + 32 * q];
#pragma unroll
                    for (int i = 0; i < R; ++i) acc[f * R + i] = fma(x[i][q], c, acc[f * R + i]);
                }
            }
            butterfly16(acc, lane);
            const int j = lane >> 1;
            if (!(lane & 1)) tw[(j % R) * tws + f0 + j / R] = acc[0];
        }
        __syncwarp();
        // ---- x' and the sums of squares ----
        double sq[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) sq[j] = 0.0;
        double mean[R];
#pragma unroll
        for (int i = 0; i < R; ++i) mean[i] = tw[i * tws] * inv_n;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const double one = Fl[32 * q];  // 1 for real columns, 0 for padding
#pragma unroll
            for (int i = 0; i < R; ++i) {
                x[i][q] = fma(-mean[i], one, x[i][q]);  // _nam.py:122
                sq[i] = fma(x[i][q], x[i][q], sq[i]);
            }
        }
        for (int rr = 0; rr < r; ++rr) {  // rank-r update X <- X - (X W^T) C^T  (_nam.py:133-135 / :146-148)
            double pr[R];
#pragma unroll
            for (int i = 0; i < R; ++i) pr[i] = -tw[i * tws + 2 + rr];
            const double *Cp = Cl + rr * LDN;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const double cv = Cp[32 * q];
#pragma unroll
                for (int i = 0; i < R; ++i) x[i][q] = fma(pr[i], cv, x[i][q]);
            }
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
#pragma unroll
            for (int i = 0; i < R; ++i) {
                sq[R + i] = fma(x[i][q], x[i][q], sq[R + i]);
                if (2 * R + i < 16) sq[2 * R + i] += x[i][q];
            }
        }
        butterfly16(sq, lane);
        {
            const int j = lane >> 1;
            if (!(lane & 1) && j < 3 * R) tw[(j % R) * tws + m1p + j / R] = sq[0];
        }
        __syncwarp();
        // ---- per-row scalars: lane i < R finishes row i ----
        double inv_std = 0.0;
        bool ok = false;
        if (lane < R && row0 + lane < a.n_rows) {
            const int64_t row = row0 + lane;
            const double *t = tw + lane * tws;
            const double mu = t[0] * inv_n;
            bool keep = true;
            if (a.row_keep) keep = a.row_keep[row] != 0;
            else if (a.qc_kurt) keep = a.qc_kurt[row] < thr;  // NaN -> dropped, _nam.py:96
            ok = keep && !(t[m1p] == 0.0);  // variance of the selected samples == 0, _association.py:182-185
            double kurt = nan("");
            if (nbk > 0 && ok) {  // _nam.py:78-82 on the residualised row
                double mm = 0.0;
                for (int b = 0; b < nbk; ++b) {
                    double vb = t[2 + r + b] - mu;
                    for (int rr = 0; rr < r; ++rr) vb = fma(-cbar[b * r + rr], t[2 + rr], vb);
                    mm += vb;
                }
                mm /= nbk;
                double m2 = 0.0, m4 = 0.0;
                for (int b = 0; b < nbk; ++b) {
                    double vb = t[2 + r + b] - mu;
                    for (int rr = 0; rr < r; ++rr) vb = fma(-cbar[b * r + rr], t[2 + rr], vb);
                    const double d2 = (vb - mm) * (vb - mm);
                    m2 += d2;
                    m4 += d2 * d2;
                }
                kurt = kurtosis_from_moments(mm, m2 / nbk, m4 / nbk, false);
            }
            if (a.kurt) a.kurt[row] = kurt;
            if (a.qc_out) {  // the QC statistic of the raw row (_nam.py:78-82): kurtosis across its batch means
                double mm = 0.0;
                for (int b = 0; b < nbk; ++b) mm += t[2 + r + b];
                mm /= nbk;
                double m2 = 0.0, m4 = 0.0;
                for (int b = 0; b < nbk; ++b) {
                    const double d2 = (t[2 + r + b] - mm) * (t[2 + r + b] - mm);
                    m2 += d2;
                    m4 += d2 * d2;
                }
                a.qc_out[row] = kurtosis_from_moments(mm, m2 / nbk, m4 / nbk, false);
            }
            // ddof=1 standardisation (_nam.py:159; pandas std is taken around the mean of x')
            const double s1 = t[m1p + 2] * inv_n;
            inv_std = rsqrt((t[m1p + 1] - dn * s1 * s1) * inv_nm1);
            double d = t[1] - mu * cy[r];  // x'.y = x.y - mean * sum(y) - (C^T y).p
            for (int rr = 0; rr < r; ++rr) d = fma(-cy[rr], t[2 + rr], d);
            a.ncorr[row] = ok ? d * inv_std * inv_n : 0.0;  // _association.py:77
            a.row_valid[row] = ok ? 1 : 0;
        }
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int64_t row = row0 + i;
            if (row >= a.n_rows) break;  // uniform over the warp
            const double sc = __shfl_sync(kFull, ok ? inv_std : 0.0, i);  // rows of dropped cells are zero
            float *o = a.x_out ? a.x_out + row * a.ld_x : nullptr;
            __half *ph = a.x16_hi ? static_cast<__half *>(a.x16_hi) + row * a.ld16 : nullptr;
            __half *pl = a.x16_hi ? static_cast<__half *>(a.x16_lo) + row * a.ld16 : nullptr;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int m = lane + 32 * q;
                const float v = float(x[i][q] * sc);
                // NQ <= 8 is instantiated exactly (NQ == ceil(n / 32)): only the last column group can run past the row
                if (o && ((NQ <= 8 && q < NQ - 1) || m < a.ld_x)) o[m] = v;
                if (ph && ((NQ <= 8 && q < NQ - 1) || m < a.ld16)) {  // v = hi + lo to 2^-22 (the difference is exact in fp32)
                    const __half h = __float2half_rn(v);
                    ph[m] = h;
                    pl[m] = __float2half_rn(v - __half2float(h));
                }
            }
        }
        __syncwarp();
    }
}

// Late QC decision (cna_resid_pass with qc_out): a warp per row, rows that pass cost one load.
__global__ void __launch_bounds__(256) qc_fixup_kernel(const double *__restrict__ qc, const double *__restrict__ median,
                                                       int64_t n_rows, float *x, int64_t ld_x, __half *hi, __half *lo,
                                                       int64_t ld16, double *kurt, double *ncorr, uint8_t *row_valid) {
    const double two_med = 2.0 * median[0];
    const double thr = two_med > 6.0 ? two_med : 6.0;  // Python's max(6, 2 * median): NaN -> 6
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    for (int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; row < n_rows; row += warps) {
        if (qc[row] < thr) continue;  // kept (NaN fails the comparison: dropped, _nam.py:96)
        if (x)
            for (int64_t c = lane; c < ld_x; c += 32) x[row * ld_x + c] = 0.f;
        if (hi)
            for (int64_t c = lane; c < ld16; c += 32) {
                hi[row * ld16 + c] = __float2half_rn(0.f);
                lo[row * ld16 + c] = __float2half_rn(0.f);
            }
        if (lane == 0) {
            if (kurt) kurt[row] = nan("");
            ncorr[row] = 0.0;
            row_valid[row] = 0;
        }
    }
}

}  // namespace cna

using namespace cna;

extern "C" {

int cna_qc_fixup(const double *qc, const double *median, int64_t n_rows, float *x, int64_t ld_x, void *x16_hi,
                 void *x16_lo, int64_t ld16, double *kurt, double *ncorr, uint8_t *row_valid, void *stream) {
    CNA_REQUIRE(qc && median && ncorr && row_valid && n_rows >= 0, "cna_qc_fixup: bad arguments");
    CNA_REQUIRE((x16_hi == nullptr) == (x16_lo == nullptr), "cna_qc_fixup: both planes or none");
    if (n_rows == 0) return CNA_OK;
    int64_t blocks = (n_rows + 7) / 8, cap = int64_t(num_sms()) * 16;
    qc_fixup_kernel<<<unsigned(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(
        qc, median, n_rows, x, ld_x, static_cast<__half *>(x16_hi), static_cast<__half *>(x16_lo), ld16, kurt, ncorr,
        row_valid);
    CNA_LAUNCHED("qc_fixup_kernel");
    return CNA_OK;
}

int cna_batch_kurtosis(const float *s, int64_t ld, int64_t n_rows, const double *inv_count,
                       const int32_t *seg_order, const int32_t *seg_off, int n_batches, int n_sel,
                       double *kurt, void *stream) {
    CNA_REQUIRE(n_rows >= 0 && n_batches >= 2, "cna_batch_kurtosis: need at least two batches");
    if (n_rows == 0) return CNA_OK;
    CNA_REQUIRE(n_sel > 0 && n_sel <= ld, "cna_batch_kurtosis: bad segment table (n_sel=%d)", n_sel);
    const int threads = 256, warps = threads / 32;
    size_t smem = sizeof(double) * (size_t(n_sel) + size_t(warps) * n_sel + size_t(warps) * n_batches) +
                  sizeof(int) * (2 * size_t(n_sel) + n_batches + 1);
    CNA_REQUIRE(smem <= 200 * 1024, "cna_batch_kurtosis: %zu bytes of shared memory needed", smem);
    CNA_CUDA(cudaFuncSetAttribute(batch_kurtosis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  int(smem)));
    int64_t blocks_needed = (n_rows + warps - 1) / warps;
    unsigned grid = unsigned(blocks_needed < int64_t(num_sms()) * 8 ? blocks_needed : int64_t(num_sms()) * 8);
    batch_kurtosis_kernel<<<grid, threads, smem, as_stream(stream)>>>(
        s, ld, n_rows, inv_count, seg_order, seg_off, n_batches, n_sel, kurt);
    CNA_LAUNCHED("batch_kurtosis_kernel");
    return CNA_OK;
}

int cna_resid_pass(const cna_resid_args *args, void *stream) {
    CNA_REQUIRE(args != nullptr, "cna_resid_pass: null args");
    const cna_resid_args &a = *args;
    CNA_REQUIRE(a.n_rows >= 0 && a.n >= 2 && a.n <= 1024, "cna_resid_pass: n must be in [2, 1024] (got %d)", a.n);
    CNA_REQUIRE(a.r >= 0 && (!a.x_out || (a.ld_x >= a.n && a.ld_x <= ((a.n + 31) / 32) * 32)),
                "cna_resid_pass: bad r/ld_x (r=%d ld_x=%lld n=%d)", a.r, (long long)a.ld_x, a.n);
    CNA_REQUIRE(a.s && a.inv_count && a.colmap && a.y && (a.x_out || a.x16_hi) && a.ncorr && a.row_valid,
                "cna_resid_pass: null pointer");
    CNA_REQUIRE(!a.x16_hi || (a.x16_lo && a.ld16 % 16 == 0 && a.ld16 >= a.n && a.ld16 <= ((a.n + 31) / 32) * 32),
                "cna_resid_pass: bad fp16 planes (ld16=%lld n=%d)", (long long)a.ld16, a.n);
    CNA_REQUIRE(a.r == 0 || (a.C && a.Wt), "cna_resid_pass: C/Wt missing");
    if (a.n_rows == 0) return CNA_OK;
    const int threads = 256, warps = threads / 32;
    const bool want_kurt = a.kurt && a.n_batches > 1;
    CNA_REQUIRE(!want_kurt || (a.seg_order && a.seg_off), "cna_resid_pass: batch segments missing");
    CNA_REQUIRE(!a.qc_kurt || a.qc_median, "cna_resid_pass: qc_kurt needs qc_median");
    CNA_REQUIRE(!a.qc_out || (want_kurt && !a.qc_kurt && !a.row_keep),
                "cna_resid_pass: qc_out needs the batch segments and no other QC input");
    int nq = (a.n + 31) / 32;
    cudaStream_t st = as_stream(stream);
    {   // the linear-functional kernel whenever its tables fit in shared memory
        ResidTables tb;
        tb.nbk = want_kurt ? a.n_batches : 0;
        tb.m1 = 2 + a.r + tb.nbk;
        static const int r_env = getenv("CNA_RESID_R") ? atoi(getenv("CNA_RESID_R")) : 0;  // experiments
        const int R = (r_env == 2 && nq >= 5 && nq <= 8) ? 2 : (nq <= 8 ? 4 : (nq <= 16 ? 2 : 1));
        const int fpc = 16 / R, m1p = (tb.m1 + fpc - 1) / fpc * fpc;
        const int nqt = nq <= 8 ? nq : (nq <= 16 ? 16 : 32), ldn = 32 * nqt;
        size_t smem = sizeof(double) * (size_t(m1p + a.r + 1) * ldn + size_t(tb.nbk) * a.r + a.r + 1 +
                                        size_t(warps) * R * (m1p + 3)) + sizeof(int) * size_t(ldn);
        if (smem <= 100 * 1024) {
            int64_t blocks_needed = (a.n_rows + int64_t(warps) * R - 1) / (int64_t(warps) * R);
            // ~12 waves of CTAs: an SM that shares its issue slots with a single-CTA kernel of another stream
            // (the permutation draw, the eigensolver) then simply takes fewer of them instead of holding the
            // last wave back (measured: 1.47 -> 1.2x ms beside the draw); the table set-up is ~3 % of a CTA
            int64_t cap = int64_t(num_sms()) * 24;
            unsigned grid = unsigned(blocks_needed < cap ? blocks_needed : cap);
#define CNA_RESID_LIN(NQ, RR)                                                                                   \
    do {                                                                                                       \
        CNA_CUDA(cudaFuncSetAttribute(resid_lin_kernel<NQ, RR>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                      int(smem)));                                                             \
        resid_lin_kernel<NQ, RR><<<grid, threads, smem, st>>>(a, tb);                                          \
    } while (0)
            switch (nq <= 8 ? nq : (nq <= 16 ? 16 : 32)) {
                case 1: CNA_RESID_LIN(1, 4); break;
                case 2: CNA_RESID_LIN(2, 4); break;
                case 3: CNA_RESID_LIN(3, 4); break;
                case 4: CNA_RESID_LIN(4, 4); break;
                case 5: if (R == 2) CNA_RESID_LIN(5, 2); else CNA_RESID_LIN(5, 4); break;
                case 6: if (R == 2) CNA_RESID_LIN(6, 2); else CNA_RESID_LIN(6, 4); break;
                case 7: if (R == 2) CNA_RESID_LIN(7, 2); else CNA_RESID_LIN(7, 4); break;
                case 8: if (R == 2) CNA_RESID_LIN(8, 2); else CNA_RESID_LIN(8, 4); break;
                case 16: CNA_RESID_LIN(16, 2); break;
                default: CNA_RESID_LIN(32, 1);
            }
#undef CNA_RESID_LIN
            CNA_LAUNCHED("resid_lin_kernel");
            return CNA_OK;
        }
    }
    CNA_REQUIRE(!a.qc_out, "cna_resid_pass: qc_out is produced by the linear-functional kernel only (tables too large)");
    const int R = nq <= 8 ? 4 : (nq <= 16 ? 2 : 1);  // rows per warp (register budget: R * NQ doubles)
    size_t smem = sizeof(double) * (2 * size_t(a.r) * a.n + 2 * size_t(a.n) + size_t(warps) * R * a.r +
                                    (want_kurt ? size_t(warps) * (a.n + a.n_batches) : 0)) +
                  sizeof(int) * (2 * size_t(a.n) + a.n_batches + 2);
    CNA_REQUIRE(smem <= 200 * 1024,
                "cna_resid_pass: n=%d, r=%d needs %zu bytes of shared memory (limit 200 KiB)", a.n, a.r, smem);
    int64_t blocks_needed = (a.n_rows + int64_t(warps) * R - 1) / (int64_t(warps) * R);
    int64_t cap = int64_t(num_sms()) * 8;
    unsigned grid = unsigned(blocks_needed < cap ? blocks_needed : cap);
#define CNA_RESID(NQ, RR, EX, RT)                                                                          \
    do {                                                                                                   \
        CNA_CUDA(cudaFuncSetAttribute(resid_kernel<NQ, RR, EX, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      int(smem)));                                                         \
        resid_kernel<NQ, RR, EX, RT><<<grid, threads, smem, st>>>(a);                                      \
    } while (0)
#define CNA_RESID_EXACT(NQ)                 \
    do {                                    \
        if (a.r <= 8) CNA_RESID(NQ, 4, true, 8); \
        else CNA_RESID(NQ, 4, true, 0);     \
    } while (0)
    switch (nq <= 8 ? nq : 0) {
        case 1: CNA_RESID_EXACT(1); break;
        case 2: CNA_RESID_EXACT(2); break;
        case 3: CNA_RESID_EXACT(3); break;
        case 4: CNA_RESID_EXACT(4); break;
        case 5: CNA_RESID_EXACT(5); break;
        case 6: CNA_RESID_EXACT(6); break;
        case 7: CNA_RESID_EXACT(7); break;
        case 8: CNA_RESID_EXACT(8); break;
        default:
            if (nq <= 16) CNA_RESID(16, 2, false, 0);
            else CNA_RESID(32, 1, false, 0);
    }
#undef CNA_RESID_EXACT
#undef CNA_RESID
    CNA_LAUNCHED("resid_kernel");
    return CNA_OK;
}

}  // extern "C"
