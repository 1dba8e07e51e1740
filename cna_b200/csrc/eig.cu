// Leading eigenpairs of the n x n Gram on the device (n <= 512), one CTA, fp64 throughout.
//
// Reference: src/cna/tools/_nam.py:105  `U, svs, UT = np.linalg.svd(NAM.dot(NAM.T))` — the association
// test (_association.py:35-74) reads only U[:, :max(ks)] and svs is the eigenvalue of the symmetric PSD
// Gram, so the leading k eigenpairs are all the hot path needs.  On the host this is LAPACK dsyevr:
// ~2 ms at n = 200 and ~20 ms at n = 500, squarely on the critical path of a call (nothing can run
// before U is known except the null GEMM).  Here the same classical pipeline runs in one launch:
//   1. Householder tridiagonalisation Q^T G Q = T of the packed lower triangle (shared memory when it
//      fits, n <= ~208, otherwise an L2-resident scratch).  The rank-2 update of step j is fused with the
//      matrix-vector product of step j+1: one read-modify-write sweep over the trailing block per step,
//      in 16-row x 32-column units (a lane owns a column: column sums stay in a register, the 16 row
//      sums are reduced with one transposing butterfly);
//   2. the k largest eigenvalues of T by multisection on Sturm counts (every thread evaluates one
//      abscissa: the intervals shrink by the number of threads per eigenvalue each round; the counts
//      use the rescaled determinant recurrence, which has no division on its dependency chain);
//   3. their eigenvectors by inverse iteration on T (a thread per eigenvalue solves its partially
//      pivoted LU system; eigenvalues closer than 1e-3 |T| are re-orthogonalised as a group every
//      iteration, the criterion of LAPACK's dstein);
//   4. back-transformation by the stored reflectors, a warp per eigenvector.
// All reductions are performed in a fixed order: the result is deterministic.
#include "common.cuh"

namespace cna {
namespace eig {

constexpr int kThreads = 512, kWarps = kThreads / 32;
constexpr double kEps = 2.220446049250313e-16, kSafeMin = 2.2250738585072014e-308;
constexpr int kInvitArrays = 6;  // 1/pivot, two super-diagonals, multiplier, interchange flag, iterate

struct Args {
    const double *G;   // [n x ldg] symmetric up to rounding: (G + G^T) / 2 is decomposed
    int64_t ldg;
    int n, k;
    double *packed;    // global scratch, n (n + 1) / 2 doubles (the matrix lives here when it does not fit in smem)
    double *refl;      // global scratch, n * n doubles: reflector j in row j (coalesced for the back-transformation)
    double *work;      // global scratch, 6 n k doubles (inverse iteration, when it does not fit in smem)
    double *w_out;     // [k] eigenvalues, descending
    double *ut_out;    // [k x n] row c = unit eigenvector of the c-th largest eigenvalue
    double *de_out;    // optional [2 n + 4]: diagonal and off-diagonal of T, SM clocks per phase (diagnostics / tests)
    int invit_in_smem;
};

// sum over the CTA; every thread gets the total (fixed order: deterministic)
__device__ __forceinline__ double block_sum(double v, double *red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    static_assert(kWarps == 16, "the tree below is written for 16 warps");
    const double t = (((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]))) +
                     (((red[8] + red[9]) + (red[10] + red[11])) + ((red[12] + red[13]) + (red[14] + red[15])));
    __syncthreads();
    return t;
}
__device__ __forceinline__ double block_max(double v, double *red) {
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = red[0];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) t = fmax(t, red[w]);
    __syncthreads();
    return t;
}

__device__ __forceinline__ int64_t tri(int i) { return int64_t(i) * (i + 1) / 2; }

// Reduce v[0..16) across the warp: afterwards lane L holds the warp total of v[L >> 1] in v[0]
// (15 shuffle steps instead of 16 x 5).
template <int HALF>
__device__ __forceinline__ void bfly_step(double (&v)[16], int lane) {
    constexpr int BIT = 2 * HALF;
    const bool upper = (lane & BIT) != 0;
#pragma unroll
    for (int j = 0; j < HALF; ++j) {
        const double send = upper ? v[j] : v[j + HALF];
        const double keep = upper ? v[j + HALF] : v[j];
        v[j] = keep + __shfl_xor_sync(kFull, send, BIT);
    }
}
__device__ __forceinline__ void bfly16(double (&v)[16], int lane) {
    bfly_step<8>(v, lane);
    bfly_step<4>(v, lane);
    bfly_step<2>(v, lane);
    bfly_step<1>(v, lane);
    v[0] += __shfl_xor_sync(kFull, v[0], 1);
}

// Number of eigenvalues of the (scaled: |d|, |e| <= 1) tridiagonal matrix smaller than x: sign changes of
// the leading principal minors p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2}.  Signs are compared on the
// high words (integer pipe) and the pair is rescaled by an exact power of two every fourth step (a
// minor changes by at most a factor ~3 upwards and, in double precision, ~1e-16 downwards per step).
__device__ __forceinline__ int sturm_count(const double *__restrict__ d, const double *__restrict__ e2, int n, double x) {
    double pm = 1.0, p = d[0] - x;
    int c = __double2hiint(p) < 0 ? 1 : 0;
    for (int i = 1; i < n; ++i) {
        const double pn = fma(d[i] - x, p, -e2[i - 1] * pm);
        c += (__double2hiint(pn) ^ __double2hiint(p)) < 0 ? 1 : 0;
        pm = p;
        p = pn;
        if ((i & 3) == 0) {
            const int ex = (__double2hiint(p) >> 20) & 0x7ff;  // biased exponent
            if (ex > 1023 + 200) {
                p *= 0x1p-200;
                pm *= 0x1p-200;
            } else if (ex < 1023 - 200) {
                p *= 0x1p+200;
                pm *= 0x1p+200;
            }
        }
    }
    return c;
}

template <int NQ, bool SMEM>
__global__ void __launch_bounds__(kThreads, 1) sym_eig_top_kernel(Args a) {
    extern __shared__ double sm[];
    const int n = a.n, k = a.k;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Storage is padded to np = the next multiple of 16 (zero rows / entries): every 16-row unit of the sweep is
    // complete, and the padding stays exactly zero through every update.
    const int np = (n + 15) & ~15;
    const int HB = np >> 4, NB = (np + 31) >> 5;  // 16-row half-blocks, 32-column blocks
    // ---- shared-memory carve-up ----
    double *v = sm;                 // [np] pending reflector
    double *w = v + np;             // [np] pending update vector
    double *vn = w + np;            // [np] reflector under construction (ping-pong with v)
    double *raw = vn + np;          // [np] column j of the updated matrix, unscaled
    double *e2 = raw + np;          // [np] squares of the scaled off-diagonal
    double *dS = e2 + np;           // [np] diagonal of T
    double *eS = dS + np;           // [np] off-diagonal of T (n - 1 used)
    double *tauS = eS + np;         // [np]
    double *redA = tauS + np;       // [kWarps]
    double *redB = redA + kWarps;   // [kWarps]
    double *lamS = redB + kWarps;   // [k] eigenvalues (ascending)
    double *loS = lamS + k;         // [k]
    double *hiS = loS + k;          // [k]
    int *group = reinterpret_cast<int *>(hiS + k);   // [k] first member of the cluster of an eigenvalue
    double *scratch = reinterpret_cast<double *>(group + ((k + 1) & ~1));
    double *rowpart = scratch;                       // [NB x np] row sums per column block
    double *colpart = rowpart + size_t(NB) * np;     // [HB x np] column sums per half-block
    double *A = SMEM ? colpart + size_t(HB) * np : a.packed;  // packed lower triangle of the padded matrix, row-major
    double *red = redA;

    const long long clk0 = clock64();
    // ---- load (G + G^T) / 2 into the packed lower triangle ----
    for (int i = warp; i < np; i += kWarps)
        for (int l = lane; l <= i; l += 32)
            A[tri(i) + l] = i < n ? 0.5 * (a.G[int64_t(i) * a.ldg + l] + a.G[int64_t(l) * a.ldg + i]) : 0.0;
    for (int i = tid; i < 4 * np; i += kThreads) v[i] = 0.0;  // v, w, vn, raw
    __syncthreads();

    // ---- 1. Householder tridiagonalisation (LAPACK dsytd2 'L' arithmetic, update fused with the next product) ----
    // Thread i owns element i of every vector (np <= kThreads), so four barriers separate the phases of a
    // step: column + norm | reflector | sweep | dot.  The pending reflector v has v[j] = 1 at its first index j
    // and the matching w[j] is carried in a register by every thread (w_first).
    bool pending = false;
    double w_first = 0.0;
    for (int j = 0; j + 2 < n; ++j) {
        const int s = j + 1, hb0 = s >> 4, bl0 = s >> 5;
        // column j with the pending update applied -> raw[j .. np-1]; sum of squares below the sub-diagonal
        {
            double sq = 0.0;
            if (tid >= j && tid < np) {
                double x = A[tri(tid) + j];
                if (pending) {
                    x = fma(-v[tid], w_first, x);
                    x -= w[tid];  // v[j] = 1
                }
                raw[tid] = x;
                if (tid >= j + 2) sq = x * x;
            }
            sq = warp_sum(sq);
            if (lane == 0) redA[warp] = sq;
        }
        __syncthreads();
        const double ss = (((redA[0] + redA[1]) + (redA[2] + redA[3])) + ((redA[4] + redA[5]) + (redA[6] + redA[7]))) +
                          (((redA[8] + redA[9]) + (redA[10] + redA[11])) + ((redA[12] + redA[13]) + (redA[14] + redA[15])));
        const double alpha = raw[s], diag = raw[j];
        double beta = alpha, t = 0.0, scale = 0.0;
        if (ss > 0.0) {  // dlarfg
            beta = -copysign(sqrt(fma(alpha, alpha, ss)), alpha);
            t = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        double vn_own = 0.0;  // element tid of the new reflector
        if (tid >= s && tid < np) {
            vn_own = (tid == s) ? 1.0 : raw[tid] * scale;
            vn[tid] = vn_own;
            if (tid < n) a.refl[size_t(j) * n + tid] = vn_own;  // kept for the back-transformation
        }
        if (tid == 0) {
            dS[j] = diag;
            eS[j] = beta;
            tauS[j] = t;
        }
        __syncthreads();
        // fused sweep over the trailing block rows / columns s .. np-1 in 16 x 32 units: apply the pending
        // rank-2 update, accumulate p = A22 . vn (row part per row, column part per lane)
        {
            int rem = warp;
            for (int hb = hb0; hb < HB; ++hb) {
                const int cnt = min((16 * hb + 15) >> 5, NB - 1) - bl0 + 1;
                for (; rem < cnt; rem += kWarps) {
                    const int bl = bl0 + rem, l = 32 * bl + lane, i0 = 16 * hb;
                    const bool lin = l >= s && l < np;
                    const double vl = (lin && pending) ? v[l] : 0.0, wl = (lin && pending) ? w[l] : 0.0;
                    const double nl = lin ? vn[l] : 0.0;
                    double *Ap = A + tri(i0) + l;  // element (i0 + r, l) at Ap[r * i0 + r (r + 1) / 2]
                    double x[16];
                    double c0 = 0.0, c1 = 0.0;
                    if (i0 >= 32 * bl + 32 && i0 >= s) {
                        // every row of the unit lies strictly below every column and inside the trailing block:
                        // no per-element predicates (lanes left of column s carry vl = wl = nl = 0: their stale
                        // entries pass through unchanged and their sums are not stored)
#pragma unroll
                        for (int r = 0; r < 16; ++r) x[r] = Ap[r * i0 + r * (r + 1) / 2];
#pragma unroll
                        for (int r = 0; r < 16; ++r) {
                            const int i = i0 + r;
                            if (pending) {
                                x[r] = fma(-v[i], wl, x[r]);
                                x[r] = fma(-w[i], vl, x[r]);
                            }
                            if (r & 1) c1 = fma(x[r], vn[i], c1);
                            else c0 = fma(x[r], vn[i], c0);
                        }
                        if (pending) {
#pragma unroll
                            for (int r = 0; r < 16; ++r) Ap[r * i0 + r * (r + 1) / 2] = x[r];
                        }
                    } else {
                        // a unit on the diagonal or on the upper edge of the trailing block: this lane owns the
                        // elements of rows r >= r_ok (on or below the diagonal, inside the block) and adds to its
                        // column sum from r_c
                        const int r_lo = max(0, s - i0);  // uniform
                        const int r_ok = lin ? max(l - i0, r_lo) : 16, r_c = lin ? max(l - i0 + 1, r_lo) : 16;
#pragma unroll
                        for (int r = 0; r < 16; ++r) x[r] = (r >= r_ok) ? Ap[r * i0 + r * (r + 1) / 2] : 0.0;
#pragma unroll
                        for (int r = 0; r < 16; ++r) {
                            const int i = i0 + r;
                            if (pending) {
                                double y = fma(-v[i], wl, x[r]);
                                y = fma(-w[i], vl, y);
                                x[r] = (r >= r_ok) ? y : 0.0;
                            }
                            const double xc = (r >= r_c) ? x[r] : 0.0;
                            if (r & 1) c1 = fma(xc, vn[i], c1);
                            else c0 = fma(xc, vn[i], c0);
                        }
                        if (pending) {
#pragma unroll
                            for (int r = 0; r < 16; ++r)
                                if (r >= r_ok) Ap[r * i0 + r * (r + 1) / 2] = x[r];
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 16; ++r) x[r] *= nl;
                    bfly16(x, lane);
                    if (!(lane & 1)) {
                        const int i = i0 + (lane >> 1);
                        if (i >= s) rowpart[bl * np + i] = x[0];
                    }
                    if (lin) colpart[hb * np + l] = c0 + c1;
                }
                rem -= cnt;
            }
        }
        __syncthreads();
        // p = tau A22 vn ; w = p - (tau / 2) (p . vn) vn ; every thread also forms the entry at index s
        double p_own = 0.0, p_first = 0.0;
        {
            const int l = (tid >= s && tid < n) ? tid : s;  // idle threads shadow the first entry
            double q0 = 0.0, q1 = 0.0, f0 = 0.0, f1 = 0.0;
            const int blmax = l >> 5;
            for (int bl = bl0; bl <= blmax; ++bl) q0 += rowpart[bl * np + l];
            f0 = rowpart[bl0 * np + s];
            int hb = max(hb0, 2 * blmax);
            for (; hb + 1 < HB; hb += 2) {
                q0 += colpart[hb * np + l];
                q1 += colpart[(hb + 1) * np + l];
            }
            if (hb < HB) q0 += colpart[hb * np + l];
            hb = max(hb0, 2 * bl0);
            for (; hb + 1 < HB; hb += 2) {
                f0 += colpart[hb * np + s];
                f1 += colpart[(hb + 1) * np + s];
            }
            if (hb < HB) f0 += colpart[hb * np + s];
            p_own = (q0 + q1) * t;
            p_first = (f0 + f1) * t;
            double dot = (tid >= s && tid < n) ? p_own * vn_own : 0.0;
            dot = warp_sum(dot);
            if (lane == 0) redB[warp] = dot;
        }
        __syncthreads();
        {
            const double dot = (((redB[0] + redB[1]) + (redB[2] + redB[3])) + ((redB[4] + redB[5]) + (redB[6] + redB[7]))) +
                               (((redB[8] + redB[9]) + (redB[10] + redB[11])) + ((redB[12] + redB[13]) + (redB[14] + redB[15])));
            const double al = -0.5 * t * dot;
            if (tid >= s && tid < n) w[tid] = fma(al, vn_own, p_own);
            w_first = p_first + al;  // vn[s] = 1
        }
        double *tmp = v;
        v = vn;
        vn = tmp;
        pending = true;
        // (no barrier: the next column phase touches only this thread's own entries and the finished matrix)
    }
    __syncthreads();
    if (tid == 0) {  // trailing 2 x 2 (or the whole matrix when n <= 2)
        if (n == 1) {
            dS[0] = A[0];
        } else {
            const int p = n - 2, q = n - 1;
            double app = A[tri(p) + p], aqp = A[tri(q) + p], aqq = A[tri(q) + q];
            if (pending) {
                app -= 2.0 * v[p] * w[p];
                aqp -= v[q] * w[p] + w[q] * v[p];
                aqq -= 2.0 * v[q] * w[q];
            }
            dS[p] = app;
            dS[q] = aqq;
            eS[p] = aqp;
        }
        eS[n - 1] = 0.0;
    }
    __syncthreads();
    if (a.de_out)
        for (int i = tid; i < n; i += kThreads) {
            a.de_out[i] = dS[i];
            a.de_out[n + i] = eS[i];
        }

    const long long clk1 = clock64();
    // ---- 2. the k largest eigenvalues of T: multisection on Sturm counts of T / |T| ----
    double gl = 1e300, gu = -1e300;
    for (int i = tid; i < n; i += kThreads) {
        const double r = (i > 0 ? fabs(eS[i - 1]) : 0.0) + (i + 1 < n ? fabs(eS[i]) : 0.0);
        gl = fmin(gl, dS[i] - r);
        gu = fmax(gu, dS[i] + r);
    }
    gl = -block_max(-gl, red);
    gu = block_max(gu, red);
    const double tnorm = fmax(fmax(fabs(gl), fabs(gu)), kSafeMin);
    const double inv_norm = 1.0 / tnorm;
    double *dsc = vn;  // scaled diagonal (the reflector buffers are free now)
    for (int i = tid; i < n; i += kThreads) {
        dsc[i] = dS[i] * inv_norm;
        const double es = eS[i] * inv_norm;
        e2[i] = es * es;
    }
    gl = gl * inv_norm - 2.1 * kEps * n;
    gu = gu * inv_norm + 2.1 * kEps * n;
    for (int c = tid; c < k; c += kThreads) {
        loS[c] = gl;
        hiS[c] = gu;
    }
    __syncthreads();
    {
        int *cnt = reinterpret_cast<int *>(scratch);  // [kThreads]
        for (int c0 = 0; c0 < k; c0 += kThreads / 4) {  // at least 4 abscissae per eigenvalue and round
            const int kc = min(k - c0, kThreads / 4);
            const int T = kThreads / kc;
            const int c = c0 + tid / T, tt = tid % T;
            const bool mine = tid / T < kc;
            const int idx = n - k + c;  // ascending index of this eigenvalue
            for (int round = 0; round < 200; ++round) {
                double lo = 0.0, hi = 0.0, x = 0.0;
                bool active = false;
                int my = -1;
                if (mine) {
                    lo = loS[c];
                    hi = hiS[c];
                    active = (hi - lo) > 2.0 * kEps * fmax(fabs(lo), fabs(hi)) + 0.25 * kEps;  // dstebz: reltol + abstol (eps |T| / 4)
                    if (active) {
                        x = lo + (hi - lo) * (double(tt + 1) / double(T + 1));
                        my = sturm_count(dsc, e2, n, x);
                    }
                }
                cnt[tid] = my;
                if (!__syncthreads_or(active ? 1 : 0)) break;
                if (active) {
                    // the largest abscissa with count <= idx becomes lo, the smallest with count > idx becomes hi
                    if (my <= idx && (tt == T - 1 || cnt[tid + 1] > idx)) loS[c] = x;
                    if (my > idx && (tt == 0 || cnt[tid - 1] <= idx)) hiS[c] = x;
                }
                __syncthreads();
            }
            __syncthreads();
        }
    }
    for (int c = tid; c < k; c += kThreads) lamS[c] = 0.5 * (loS[c] + hiS[c]) * tnorm;
    __syncthreads();

    const long long clk2 = clock64();
    // ---- 3. eigenvectors of T by inverse iteration (LAPACK dstein's scheme) ----
    if (tid == 0) {
        group[0] = 0;
        for (int c = 1; c < k; ++c) {
            // dstein: eigenvalues closer than 10 eps |lambda| are pulled apart
            if (lamS[c] - lamS[c - 1] < 10.0 * kEps * fabs(lamS[c])) lamS[c] = lamS[c - 1] + 10.0 * kEps * fabs(lamS[c]);
            group[c] = (fabs(lamS[c] - lamS[c - 1]) < 1e-3 * tnorm) ? group[c - 1] : c;
        }
    }
    __syncthreads();
    // per-eigenvalue arrays, element i of eigenvalue c at [(array * n + i) * k + c]: in the shared memory the
    // matrix has left behind when they fit, otherwise in global scratch
    double *fu = a.invit_in_smem ? scratch : a.work;
    double *fv = fu + size_t(n) * k, *fw = fv + size_t(n) * k, *fl = fw + size_t(n) * k, *fs = fl + size_t(n) * k,
           *Y = fs + size_t(n) * k;
    const double eps3 = kEps * tnorm;
    if (tid < k) {  // partially pivoted LU of T - lambda I
        const int c = tid;
        const double lam = lamS[c];
        double aa = dS[0] - lam, bb = n > 1 ? eS[0] : 0.0;
        for (int i = 0; i + 1 < n; ++i) {
            const double cc = eS[i], dn = dS[i + 1] - lam, en = (i + 2 < n) ? eS[i + 1] : 0.0;
            const bool swap = fabs(cc) > fabs(aa);  // interchange rows i and i + 1
            if (!swap && aa == 0.0) aa = eps3;
            const double rp = 1.0 / (swap ? cc : aa), m = (swap ? aa : cc) * rp;
            fu[size_t(i) * k + c] = rp;
            fv[size_t(i) * k + c] = swap ? dn : bb;
            fw[size_t(i) * k + c] = swap ? en : 0.0;
            fl[size_t(i) * k + c] = m;
            fs[size_t(i) * k + c] = swap ? 1.0 : 0.0;
            if (swap) {
                aa = fma(-m, dn, bb);
                bb = -m * en;
            } else {
                aa = fma(-m, bb, dn);
                bb = en;
            }
        }
        if (aa == 0.0) aa = eps3;
        fu[size_t(n - 1) * k + c] = 1.0 / aa;
    }
    for (int t = tid; t < n * k; t += kThreads) {  // deterministic start vectors in (-1, 1)
        uint32_t h = uint32_t(t) * 2654435761u + 0x9e3779b9u;
        h ^= h >> 16;
        h *= 0x85ebca6bu;
        h ^= h >> 13;
        h *= 0xc2b2ae35u;
        h ^= h >> 16;
        Y[t] = double(h) * (2.0 / 4294967296.0) - 1.0 + 1.0 / 4294967296.0;
    }
    __syncthreads();
    for (int iter = 0; iter < 3; ++iter) {
        if (tid < k) {
            const int c = tid;
            double *y = Y + c;
            double yi = y[0];
            for (int i = 0; i + 1 < n; ++i) {  // forward substitution with the recorded interchanges
                double yn = y[size_t(i + 1) * k];
                if (fs[size_t(i) * k + c] != 0.0) {
                    const double tsw = yi;
                    yi = yn;
                    yn = tsw;
                }
                y[size_t(i) * k] = yi;
                yi = fma(-fl[size_t(i) * k + c], yi, yn);
            }
            double y1 = yi * fu[size_t(n - 1) * k + c], y2 = 0.0;  // back substitution
            y[size_t(n - 1) * k] = y1;
            for (int i = n - 2; i >= 0; --i) {
                const double x = (y[size_t(i) * k] - fv[size_t(i) * k + c] * y1 - fw[size_t(i) * k + c] * y2) *
                                 fu[size_t(i) * k + c];
                y[size_t(i) * k] = x;
                y2 = y1;
                y1 = x;
            }
        }
        __syncthreads();
        // re-orthogonalise inside clusters (in order) and normalise: a warp per cluster
        int ordinal = 0;
        for (int c = 0; c < k; ++c) {
            if (group[c] != c) continue;
            if ((ordinal++ % kWarps) != warp) continue;
            for (int m = c; m < k && group[m] == c; ++m) {
                for (int m2 = c; m2 < m; ++m2) {
                    double dt = 0.0;
                    for (int i = lane; i < n; i += 32) dt += Y[size_t(i) * k + m] * Y[size_t(i) * k + m2];
                    dt = warp_sum(dt);
                    for (int i = lane; i < n; i += 32) Y[size_t(i) * k + m] -= dt * Y[size_t(i) * k + m2];
                    __syncwarp();
                }
                double mx = 0.0;
                for (int i = lane; i < n; i += 32) mx = fmax(mx, fabs(Y[size_t(i) * k + m]));
                mx = warp_max(mx);
                const double inv_mx = mx > 0.0 ? 1.0 / mx : 1.0;  // scale first: the iterate grows by ~1/eps per solve
                double nn = 0.0;
                for (int i = lane; i < n; i += 32) {
                    const double x = Y[size_t(i) * k + m] * inv_mx;
                    nn += x * x;
                }
                nn = warp_sum(nn);
                const double sc = nn > 0.0 ? inv_mx / sqrt(nn) : 0.0;
                for (int i = lane; i < n; i += 32) Y[size_t(i) * k + m] *= sc;
                __syncwarp();
            }
        }
        __syncthreads();
    }

    const long long clk3 = clock64();
    // ---- 4. back-transformation u = H_0 H_1 ... H_{n-3} z, a warp per eigenvector ----
    for (int c = warp; c < k; c += kWarps) {
        double u[NQ], hv[NQ], hnext[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int i = lane + 32 * q;
            u[q] = i < n ? Y[size_t(i) * k + c] : 0.0;
            hnext[q] = (n >= 3 && i < n && i > n - 3) ? a.refl[size_t(n - 3) * n + i] : 0.0;
        }
        for (int j = n - 3; j >= 0; --j) {
            double dt = 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int i = lane + 32 * q;
                hv[q] = hnext[q];
                hnext[q] = (j > 0 && i < n && i > j - 1) ? a.refl[size_t(j - 1) * n + i] : 0.0;  // prefetch
                dt = fma(hv[q], u[q], dt);
            }
            dt = warp_sum(dt) * tauS[j];
#pragma unroll
            for (int q = 0; q < NQ; ++q) u[q] = fma(-dt, hv[q], u[q]);
        }
        const int out = k - 1 - c;  // descending order
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int i = lane + 32 * q;
            if (i < n) a.ut_out[size_t(out) * n + i] = u[q];
        }
        if (lane == 0) a.w_out[out] = lamS[c];
    }
    if (a.de_out && tid == 0) {  // SM clocks spent in the four phases (diagnostics)
        const long long clk4 = clock64();
        a.de_out[2 * n + 0] = double(clk1 - clk0);
        a.de_out[2 * n + 1] = double(clk2 - clk1);
        a.de_out[2 * n + 2] = double(clk3 - clk2);
        a.de_out[2 * n + 3] = double(clk4 - clk3);
    }
}

struct Layout {
    size_t fixed;     // bytes before `scratch`
    size_t partial;   // rowpart + colpart
    size_t matrix;    // packed lower triangle
};
static Layout layout(int n, int k) {
    const int np = (n + 15) & ~15, HB = np / 16, NB = (np + 31) / 32;
    Layout L;
    L.fixed = sizeof(double) * (size_t(8) * np + 2 * kWarps + size_t(3) * k) + sizeof(int) * size_t((k + 1) & ~1);
    L.partial = sizeof(double) * size_t(HB + NB) * np;
    if (L.partial < sizeof(int) * kThreads) L.partial = sizeof(int) * kThreads;  // the Sturm-count table lives there too
    L.matrix = sizeof(double) * size_t(np) * (np + 1) / 2;
    return L;
}

}  // namespace eig
}  // namespace cna

using namespace cna;

extern "C" int64_t cna_sym_eig_workspace(int n, int k) {
    if (n < 1 || k < 1) return 0;
    const int64_t np = (n + 15) & ~15;
    return int64_t(sizeof(double)) * (np * (np + 1) / 2 + int64_t(n) * n + int64_t(eig::kInvitArrays) * n * k);
}

extern "C" int cna_sym_eig_top(const double *G, int64_t ldg, int n, int k, double *w_out, double *ut_out,
                               double *de_out, void *workspace, int64_t workspace_bytes, void *stream) {
    CNA_REQUIRE(G && w_out && ut_out && workspace, "cna_sym_eig_top: null argument");
    CNA_REQUIRE(n >= 1 && n <= 512 && k >= 1 && k <= n && ldg >= n, "cna_sym_eig_top: bad shape (n=%d k=%d ldg=%lld)",
                n, k, (long long)ldg);
    CNA_REQUIRE(workspace_bytes >= cna_sym_eig_workspace(n, k), "cna_sym_eig_top: workspace too small");
    eig::Args a;
    a.G = G;
    a.ldg = ldg;
    a.n = n;
    a.k = k;
    a.packed = static_cast<double *>(workspace);
    const size_t np = size_t((n + 15) & ~15);
    a.refl = a.packed + np * (np + 1) / 2;
    a.work = a.refl + size_t(n) * n;
    a.w_out = w_out;
    a.ut_out = ut_out;
    a.de_out = de_out;
    cudaStream_t st = as_stream(stream);
    const size_t limit = 227 * 1024;
    const eig::Layout L = eig::layout(n, k);
    const size_t without = L.fixed + L.partial, with = without + L.matrix;
    CNA_REQUIRE(without <= limit, "cna_sym_eig_top: shared memory (%zu bytes) exceeds the limit", without);
    const bool in_smem = with <= limit;
    const size_t bytes = in_smem ? with : without;
    // the inverse iteration reuses everything from `scratch` on (partial tables + matrix) when it fits
    a.invit_in_smem = sizeof(double) * size_t(eig::kInvitArrays) * n * k <= bytes - L.fixed ? 1 : 0;
#define CNA_EIG(NQ, SM)                                                                                     \
    do {                                                                                                    \
        CNA_CUDA(cudaFuncSetAttribute(eig::sym_eig_top_kernel<NQ, SM>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      int(bytes)));                                                         \
        eig::sym_eig_top_kernel<NQ, SM><<<1, eig::kThreads, bytes, st>>>(a);                                \
    } while (0)
    if (in_smem && n <= 128) CNA_EIG(4, true);
    else if (in_smem) CNA_EIG(7, true);
    else if (n <= 256) CNA_EIG(8, false);
    else CNA_EIG(16, false);
#undef CNA_EIG
    CNA_LAUNCHED("sym_eig_top_kernel");
    return CNA_OK;
}
