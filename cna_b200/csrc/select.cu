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

constexpr int kSelBits = 16, kSelBins = 1 << kSelBits, kSelPasses = 64 / kSelBits;
constexpr unsigned long long kSkipBits = CNA_MEDIAN_SKIP_BITS;

struct SelectState {
    unsigned long long prefix;    // bits decided so far (high digits of the key of rank `rank`)
    unsigned long long rank;      // rank still to be resolved inside the current prefix
    unsigned long long n_valid;   // finite or infinite, non-NaN, not masked
    unsigned long long n_nan;
    unsigned long long max_less;  // largest key strictly below the selected one (0 if none)
    unsigned int blocks_done;
    unsigned int n_less;          // elements strictly below the selected key
};

__device__ __forceinline__ unsigned long long sortable(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double unsortable(unsigned long long k) {
    unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)u);
}

// PASS p looks at digit p (from the top) of every key whose higher digits equal the prefix found so
// far.  The last block to finish scans the histogram, extends the prefix and clears the histogram for
// the next pass.  Pass 0 also counts the population and the NaNs and turns "the upper median" into a
// rank.
template <int PASS>
__global__ void __launch_bounds__(512)
select_pass_kernel(const double *__restrict__ v, const uint8_t *__restrict__ valid, int64_t n,
                   unsigned int *__restrict__ hist, SelectState *__restrict__ st) {
    const int lane = threadIdx.x & 31;
    const int shift = 64 - kSelBits * (PASS + 1);
    const unsigned long long prefix = PASS == 0 ? 0ull : st->prefix;
    unsigned int my_nan = 0;
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    const int64_t n_pad = (n + 31) / 32 * 32;  // whole warps stay in the loop: match_any needs them
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_pad; i += stride) {
        unsigned int bin = 0x80000000u | lane;  // no partner, not counted
        if (i < n && (!valid || valid[i])) {
            const double x = v[i];
            const unsigned long long raw = (unsigned long long)__double_as_longlong(x);
            if (raw == kSkipBits) {
            } else if (x != x) {
                if (PASS == 0) ++my_nan;
            } else {
                const unsigned long long key = sortable(x);
                if (PASS == 0 || (key >> (shift + kSelBits)) == (prefix >> (shift + kSelBits)))
                    bin = (unsigned int)((key >> shift) & (kSelBins - 1));
            }
        }
        // values cluster (kurtoses share their exponent): one atomic per distinct digit per warp
        const unsigned int peers = __match_any_sync(kFull, bin);
        if (!(bin & 0x80000000u) && lane == __ffs(peers) - 1) atomicAdd(hist + bin, (unsigned int)__popc(peers));
    }
    if (PASS == 0) {
        my_nan = __reduce_add_sync(kFull, my_nan);
        if (lane == 0 && my_nan) atomicAdd(&st->n_nan, (unsigned long long)my_nan);
    }
    __shared__ bool last;
    __shared__ unsigned long long part[512];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&st->blocks_done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    // ---- the last block: find the digit that holds the wanted rank ----
    constexpr int kPer = kSelBins / 512;
    unsigned long long sum = 0;
    for (int b = 0; b < kPer; ++b) sum += __ldcg(hist + threadIdx.x * kPer + b);
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long total = 0;
        for (int t = 0; t < 512; ++t) total += part[t];
        unsigned long long rank = st->rank;
        if (PASS == 0) {
            st->n_valid = total;
            rank = total / 2;  // upper median; the lower one is resolved by select_finish_kernel
        }
        unsigned long long before = 0;
        int t = 0;
        while (t < 511 && before + part[t] <= rank) before += part[t++];
        int b = t * kPer;
        for (;; ++b) {
            const unsigned long long c = __ldcg(hist + b);
            if (before + c > rank || b == t * kPer + kPer - 1) break;
            before += c;
        }
        st->prefix = prefix | ((unsigned long long)b << shift);
        st->rank = rank - before;
        st->blocks_done = 0;
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kSelBins; b += 512) hist[b] = 0;
}

// With the selected key known: count / find the largest of the elements below it (the lower median of an
// even population is either the same value or that element), then write the result.
__global__ void __launch_bounds__(512)
select_finish_kernel(const double *__restrict__ v, const uint8_t *__restrict__ valid, int64_t n,
                     SelectState *__restrict__ st, double *__restrict__ out) {
    const unsigned long long sel = st->prefix;
    unsigned long long best = 0;
    unsigned int less = 0;
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
    less = __reduce_add_sync(kFull, less);
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(kFull, best, o);
        best = other > best ? other : best;
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
        if ((nv & 1ull) == 0 && (unsigned long long)(*((volatile unsigned int *)&st->n_less)) == nv / 2)
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
    int64_t blocks = (n + 511) / 512;
    const int64_t cap = int64_t(num_sms()) * 4;
    unsigned grid = unsigned(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
    select_pass_kernel<0><<<grid, 512, 0, s>>>(v, valid, n, hist, st);
    select_pass_kernel<1><<<grid, 512, 0, s>>>(v, valid, n, hist, st);
    select_pass_kernel<2><<<grid, 512, 0, s>>>(v, valid, n, hist, st);
    select_pass_kernel<3><<<grid, 512, 0, s>>>(v, valid, n, hist, st);
    select_finish_kernel<<<grid, 512, 0, s>>>(v, valid, n, st, out);
    CNA_LAUNCHED("select_pass_kernel");
    count_launch(4);
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
