// Order statistics and the small data-dependent scalars of the hot path, computed where the data
// is: nothing here makes the host wait.
//
//   cna_median_f64      np.median of a device vector (radix select on the sortable bit pattern of the
//                       doubles, 16 bits per pass), result left in device memory.
//                       replaces: _nam.py:59 `np.median(st.kurtosis(...))`, :94 `np.median(kurtoses)`,
//                       :153 `np.median(_batch_kurtosis(...))` (the reference sorts on the host).
//   cna_fdr_thresholds  _association.py:101-102 `maxcorr = max(abs(ncorrs).max(), 0.001)`,
//                       `np.arange(maxcorr/4, maxcorr, maxcorr/400)` and _stats.py:51 (the histogram
//                       edges) from the device-resident max |ncorr|, operation for operation what
//                       numpy's arange / fill loop does in float64 (no contraction into FMAs).
#include "common.cuh"

namespace cna {

constexpr int kSelBits = 11, kSelBins = 1 << kSelBits, kSelPasses = (64 + kSelBits - 1) / kSelBits;  // 6 passes
constexpr int kSelThreads = 1024;
constexpr unsigned long long kSkipBits = CNA_MEDIAN_SKIP_BITS;

struct SelectState {
    unsigned long long prefix;    // bits decided so far (high digits of the key of rank `rank`)
    unsigned long long rank;      // rank still to be resolved inside the current prefix
    unsigned long long n_valid;   // finite or infinite, non-NaN, not masked
    unsigned long long n_nan;
    unsigned long long max_less;  // largest key strictly below the selected one (0 if none)
    unsigned long long n_less;    // elements strictly below the selected key
    unsigned int blocks_done;
    unsigned int pad;
};

__device__ __forceinline__ unsigned long long sortable(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double unsortable(unsigned long long k) {
    unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)u);
}

// PASS p looks at digit p (from the top, 11 bits, the last one 9) of every key whose higher digits equal
// the prefix found so far.  Counts go to a shared-memory histogram per CTA (kurtoses share their exponent:
// the hot bins would serialise in L2), non-empty bins are flushed to the global one, and the last CTA to
// finish scans it, extends the prefix and clears it for the next pass.  Pass 0 also counts the
// population and the NaNs and turns "the upper median" into a rank.
template <int PASS>
__global__ void __launch_bounds__(kSelThreads)
select_pass_kernel(const double *__restrict__ v, const uint8_t *__restrict__ valid, int64_t n,
                   unsigned int *__restrict__ hist, SelectState *__restrict__ st) {
    constexpr int kHigh = PASS * kSelBits;                                   // bits already decided
    constexpr int kBits = 64 - kHigh < kSelBits ? 64 - kHigh : kSelBits;     // width of this digit
    constexpr int kShift = 64 - kHigh - kBits;
    __shared__ unsigned int h[kSelBins];
    __shared__ unsigned long long wsum[kSelThreads / 32];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b = threadIdx.x; b < kSelBins; b += kSelThreads) h[b] = 0;
    __syncthreads();
    unsigned long long prefix = 0;
    if (PASS > 0) prefix = st->prefix;
    unsigned int my_nan = 0;
    const int64_t stride = int64_t(gridDim.x) * kSelThreads;
    for (int64_t i = int64_t(blockIdx.x) * kSelThreads + threadIdx.x; i < n; i += stride) {
        if (valid && !valid[i]) continue;
        const double x = v[i];
        if (x != x) {  // NaN (counted, numpy's median is then NaN) or the skip pattern (ignored)
            if (PASS == 0 && (unsigned long long)__double_as_longlong(x) != kSkipBits) ++my_nan;
            continue;
        }
        const unsigned long long key = sortable(x);
        if (PASS > 0 && ((key ^ prefix) >> (64 - (PASS > 0 ? kHigh : 1))) != 0) continue;
        atomicAdd(h + (unsigned int)((key >> kShift) & ((1u << kBits) - 1)), 1u);
    }
    if (PASS == 0) {
        my_nan = __reduce_add_sync(kFull, my_nan);
        if (lane == 0 && my_nan) atomicAdd(&st->n_nan, (unsigned long long)my_nan);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kSelBins; b += kSelThreads)
        if (h[b]) atomicAdd(hist + b, h[b]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&st->blocks_done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    // ---- the last CTA: the bin whose cumulative count first exceeds the wanted rank (2 bins per thread) ----
    const unsigned long long c0 = __ldcg(hist + 2 * threadIdx.x), c1 = __ldcg(hist + 2 * threadIdx.x + 1);
    unsigned long long incl = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long up = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned long long base = 0, total = 0;
    for (int k = 0; k < kSelThreads / 32; ++k) {
        const unsigned long long t = wsum[k];
        if (k < warp) base += t;
        total += t;
    }
    incl += base;
    const unsigned long long excl = incl - (c0 + c1);
    unsigned long long rank = PASS == 0 ? total / 2 : st->rank;  // pass 0: the upper median
    if (PASS == 0 && threadIdx.x == 0) st->n_valid = total;
    __syncthreads();  // every thread has read st->rank before it is rewritten
    bool mine = rank < total ? (excl <= rank && rank < incl) : threadIdx.x == kSelThreads - 1;  // empty: NaN anyway
    if (mine) {
        const bool second = rank < total && rank >= excl + c0;
        const unsigned long long b = 2 * threadIdx.x + (second ? 1 : 0);
        st->prefix = prefix | (b << kShift);
        st->rank = rank < total ? rank - excl - (second ? c0 : 0) : 0;
        st->blocks_done = 0;
    }
    for (int b = threadIdx.x; b < kSelBins; b += kSelThreads) hist[b] = 0;
}

// With the selected key known: count / find the largest of the elements below it (the lower median of an
// even population is either the same value or that element), then write the result.
__global__ void __launch_bounds__(512)
select_finish_kernel(const double *__restrict__ v, const uint8_t *__restrict__ valid, int64_t n,
                     SelectState *__restrict__ st, double *__restrict__ out) {
    const unsigned long long sel = st->prefix;
    unsigned long long best = 0, less = 0;
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (valid && !valid[i]) continue;
        const double x = v[i];
        if (x != x) continue;  // NaNs and the skip pattern
        const unsigned long long key = sortable(x);
        if (key < sel) {
            ++less;
            best = key > best ? key : best;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(kFull, best, o);
        best = other > best ? other : best;
        less += __shfl_xor_sync(kFull, less, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (less) atomicAdd(&st->n_less, less);
        if (best) atomicMax(&st->max_less, best);
    }
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&st->blocks_done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last || threadIdx.x != 0) return;
    __threadfence();
    const unsigned long long nv = *((volatile unsigned long long *)&st->n_valid);
    const unsigned long long nn = *((volatile unsigned long long *)&st->n_nan);
    double med;
    if (nv == 0 || nn != 0) {
        med = nan("");  // np.median: NaN in -> NaN out; empty -> NaN
    } else {
        const double hi = unsortable(sel);
        double lo = hi;
        // even population: the lower median has rank nv/2 - 1; it is below `hi` only when no duplicate of
        // `hi` sits in front of rank nv/2
        if ((nv & 1ull) == 0 && *((volatile unsigned long long *)&st->n_less) == nv / 2)
            lo = unsortable(*((volatile unsigned long long *)&st->max_less));
        med = __dmul_rn(__dadd_rn(lo, hi), 0.5);  // np.mean of the two middle values
    }
    out[0] = med;
    out[1] = double(nv + nn);
}

__global__ void fdr_thresholds_kernel(const double *__restrict__ maxabs, int cap, double *__restrict__ thresholds,
                                      double *__restrict__ edges, int *__restrict__ n_thresholds) {
    // Python: maxcorr = max(abs(ncorrs).max(), 0.001); np.arange(maxcorr/4, maxcorr, maxcorr/400)
    const double a = maxabs[0];
    const double m = (0.001 > a) ? 0.001 : a;  // max(a, 0.001) returns a unless 0.001 > a (NaN stays)
    const double start = __ddiv_rn(m, 4.0), step = __ddiv_rn(m, 400.0);
    // numpy: length = ceil((stop - start) / step); first two values start, start + step; the rest
    // start + i * delta with delta = (start + step) - start
    double len_d = ceil(__ddiv_rn(__dadd_rn(m, -start), step));
    if (!(len_d >= 0.0)) len_d = 0.0;  // NaN / negative -> empty
    int T = len_d > double(cap) ? cap : int(len_d);
    const double next = __dadd_rn(start, step), delta = __dadd_rn(next, -start);
    for (int i = threadIdx.x; i < cap; i += blockDim.x) {
        double t = 0.0;
        if (i == 0) t = start;
        else if (i == 1) t = next;
        else t = __dadd_rn(start, __dmul_rn(double(i), delta));
        if (i >= T) t = 0.0;
        thresholds[i] = t;
        const double t2 = __dmul_rn(t, t);  // _stats.py:51: t**2 - atol - rtol * t**2
        edges[i] = i < T ? __dadd_rn(__dadd_rn(t2, -1e-8), -__dmul_rn(1e-5, t2)) : 0.0;
    }
    if (threadIdx.x == 0) n_thresholds[0] = T;
}

// _stats.py:57-59, :79-80 and _association.py:234 on the device, from the histograms the null GEMM and
// cna_obs_hist left there: tails = reverse cumulative sums, fdr_i = (sum_k tails[k, i]) / ranks_i / n_null
// (the same two float64 divisions numpy performs), pmin = running minimum of fdr that skips NaN
// (np.fmin.accumulate; Series.min() in the per-cell lookup).  One block; T <= cap <= 1024.
__global__ void __launch_bounds__(1024)
fdr_table_kernel(const unsigned long long *__restrict__ null_hist, const unsigned int *__restrict__ rank_hist,
                 const int *__restrict__ count, int cap, double n_null, double *__restrict__ fdr,
                 double *__restrict__ pmin) {
    __shared__ unsigned long long tn[1024], tr[1024];
    __shared__ double f[1024];
    const int T = min(count[0], cap), i = threadIdx.x;
    tn[i] = i < T ? null_hist[i] : 0ull;
    tr[i] = i < T ? (unsigned long long)rank_hist[i] : 0ull;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // suffix sums (Hillis-Steele)
        const unsigned long long a = i + o < 1024 ? tn[i + o] : 0ull, b = i + o < 1024 ? tr[i + o] : 0ull;
        __syncthreads();
        tn[i] += a;
        tr[i] += b;
        __syncthreads();
    }
    f[i] = i < T ? __ddiv_rn(__ddiv_rn(double((long long)tn[i]), double((long long)tr[i])), n_null) : nan("");
    __syncthreads();
    if (i < cap) fdr[i] = i < T ? f[i] : 0.0;
    if (i == 0) {  // running fmin, 300 steps
        double m = nan("");
        for (int k = 0; k < cap; ++k) {
            if (k < T) m = fmin(m, f[k]);  // fmin ignores a NaN operand
            pmin[k] = k < T ? m : 0.0;
        }
    }
}

}  // namespace cna

using namespace cna;

extern "C" {

int64_t cna_median_workspace(void) { return int64_t(sizeof(unsigned int)) * kSelBins + int64_t(sizeof(SelectState)); }

int cna_median_f64(const double *v, const uint8_t *valid, int64_t n, double *out, void *workspace,
                   int64_t workspace_bytes, void *stream) {
    CNA_REQUIRE(n >= 0 && out && workspace && workspace_bytes >= cna_median_workspace(),
                "cna_median_f64: bad arguments (n=%lld, workspace %lld bytes)", (long long)n, (long long)workspace_bytes);
    CNA_REQUIRE(n == 0 || v, "cna_median_f64: null input");
    cudaStream_t s = as_stream(stream);
    unsigned int *hist = static_cast<unsigned int *>(workspace);
    SelectState *st = reinterpret_cast<SelectState *>(hist + kSelBins);
    CNA_CUDA(cudaMemsetAsync(workspace, 0, size_t(cna_median_workspace()), s));
    int64_t blocks = (n + 4 * kSelThreads - 1) / (4 * kSelThreads);
    const int64_t cap = int64_t(num_sms());
    unsigned grid = unsigned(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
    static_assert(kSelPasses == 6, "one launch per digit below");
    select_pass_kernel<0><<<grid, kSelThreads, 0, s>>>(v, valid, n, hist, st);
    select_pass_kernel<1><<<grid, kSelThreads, 0, s>>>(v, valid, n, hist, st);
    select_pass_kernel<2><<<grid, kSelThreads, 0, s>>>(v, valid, n, hist, st);
    select_pass_kernel<3><<<grid, kSelThreads, 0, s>>>(v, valid, n, hist, st);
    select_pass_kernel<4><<<grid, kSelThreads, 0, s>>>(v, valid, n, hist, st);
    select_pass_kernel<5><<<grid, kSelThreads, 0, s>>>(v, valid, n, hist, st);
    select_finish_kernel<<<grid, 512, 0, s>>>(v, valid, n, st, out);
    CNA_LAUNCHED("select_pass_kernel");
    count_launch(kSelPasses);
    return CNA_OK;
}

int cna_fdr_table(const uint64_t *null_hist, const uint32_t *rank_hist, const int32_t *n_thresholds, int cap,
                  int n_null, double *fdr, double *prefix_min_fdr, void *stream) {
    CNA_REQUIRE(null_hist && rank_hist && n_thresholds && fdr && prefix_min_fdr && cap >= 1 && cap <= 1024 && n_null > 0,
                "cna_fdr_table: bad arguments");
    fdr_table_kernel<<<1, 1024, 0, as_stream(stream)>>>(reinterpret_cast<const unsigned long long *>(null_hist),
                                                        rank_hist, n_thresholds, cap, double(n_null), fdr,
                                                        prefix_min_fdr);
    CNA_LAUNCHED("fdr_table_kernel");
    return CNA_OK;
}

int cna_fdr_thresholds(const double *maxabs, int cap, double *thresholds, double *edges, int32_t *n_thresholds,
                       void *stream) {
    CNA_REQUIRE(maxabs && thresholds && edges && n_thresholds && cap >= 2, "cna_fdr_thresholds: bad arguments");
    fdr_thresholds_kernel<<<1, 256, 0, as_stream(stream)>>>(maxabs, cap, thresholds, edges, n_thresholds);
    CNA_LAUNCHED("fdr_thresholds_kernel");
    return CNA_OK;
}

}  // extern "C"
