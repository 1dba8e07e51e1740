// Host-side permutation drawing, bit-exact with the reference's numpy calls.
//
// Reference: src/cna/tools/_stats.py:12 and :31 draw permutations as
//     np.argsort(np.random.randn(rows, num), axis=0)
// from numpy's *legacy global* generator (MT19937 + polar Box-Muller with one cached deviate), one
// block per batch.  "Bit-exact permutation ranks" therefore pins the random stream: a device RNG can
// never reproduce it.  numpy spends ~30 ms on the 2M deviates of the 1M-cell configuration and
// another ~40 ms on the strided axis-0 argsort and index scatter, all serial, which made this the
// wall-clock critical path of association().  This file restates the generator so that
//   (1) the inherently serial part (MT19937 words, rejection test) runs alone,
//   (2) the log/sqrt transform of the accepted pairs runs on all host threads,
//   (3) the per-column argsort + scatter into the [num x n] index matrix runs on all host threads,
// and the generator state is handed back so that np.random continues exactly where the reference
// would have left it.
//
// numpy sources restated (numpy/random/src): mt19937/mt19937.c (mt19937_gen, tempering),
// mt19937.h (mt19937_next_double: (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53), and
// legacy/legacy-distributions.c (legacy_gauss).  The floating-point expressions are kept literally
// (no FMA contraction: this translation unit is built by g++ with -ffp-contract=off for baseline
// x86-64) and use the same libm log/sqrt the numpy extension resolves to.
#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/cna_b200.h"

namespace cna {
int set_error(int code, const char *fmt, ...);  // api.cu
}
#define CNA_REQUIRE(cond, ...)                                            \
    do {                                                                  \
        if (!(cond)) return cna::set_error(CNA_ERR_INVALID, __VA_ARGS__); \
    } while (0)

namespace cna {
namespace hostperm {

constexpr int kN = 624, kM = 397;
constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;

// One block of the generator, out of place: `next` = the 624 state words that follow `cur`
// (numpy's mt19937_gen, restated so that the old block is kept: every block is a possible final
// state).  The first loop reads only old words; the second reads the new words written 227 places
// earlier.
static void mt_next_block(const uint32_t *__restrict__ cur, uint32_t *__restrict__ next) {
    int i = 0;
    for (; i < kN - kM; ++i) {
        uint32_t y = (cur[i] & kUpper) | (cur[i + 1] & kLower);
        next[i] = cur[i + kM] ^ (y >> 1) ^ (-(y & 1u) & kMatrixA);
    }
    for (; i < kN - 1; ++i) {
        uint32_t y = (cur[i] & kUpper) | (cur[i + 1] & kLower);
        next[i] = next[i + (kM - kN)] ^ (y >> 1) ^ (-(y & 1u) & kMatrixA);
    }
    uint32_t y = (cur[kN - 1] & kUpper) | (next[0] & kLower);
    next[kN - 1] = next[kM - 1] ^ (y >> 1) ^ (-(y & 1u) & kMatrixA);
}

static inline uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// Attempt j of legacy_gauss's rejection loop = untempered words w[4j .. 4j+3] of the stream:
// two mt19937_next_double draws, x = 2u - 1, r2 = x1^2 + x2^2 (accepted when 0 < r2 < 1).
static inline double polar_attempt(const uint32_t *w, double &x1, double &x2) {
    int32_t a1 = int32_t(mt_temper(w[0]) >> 5), b1 = int32_t(mt_temper(w[1]) >> 6);
    int32_t a2 = int32_t(mt_temper(w[2]) >> 5), b2 = int32_t(mt_temper(w[3]) >> 6);
    double d1 = (a1 * 67108864.0 + b1) / 9007199254740992.0;
    double d2 = (a2 * 67108864.0 + b2) / 9007199254740992.0;
    x1 = 2.0 * d1 - 1.0;
    x2 = 2.0 * d2 - 1.0;
    return x1 * x1 + x2 * x2;
}

static inline double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static inline bool timing_on() {
    static const bool on = std::getenv("CNA_B200_TIMING") != nullptr;
    return on;
}
static inline int default_threads() {
    // all hardware threads but one: measured on the 16-thread B200 host the draw for 10 000 x 200
    // takes 7 ms with 16 workers and 16 ms with 8, and it is over before the GPU has finished the
    // NAM, i.e. before the caller needs its cores for the n x n SVD
    unsigned hw = std::thread::hardware_concurrency();
    int t = hw > 2 ? int(hw) - 1 : 1;
    return t > 64 ? 64 : t;
}

template <typename F>
static void parallel_for(int64_t n, int n_threads, F fn) {
    if (n_threads <= 1 || n < 2) {
        fn(0, n);
        return;
    }
    std::vector<std::thread> pool;
    int64_t per = (n + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        int64_t a = t * per, b = std::min(n, a + per);
        if (a >= b) break;
        pool.emplace_back([=] { fn(a, b); });
    }
    for (auto &th : pool) th.join();
}

// argsort of n <= 256 keys by counting: rank_i = #{j : key_j < key_i}.  Branch-free and SIMD-friendly,
// ~3x faster than std::sort on the 50-element columns of a 4-batch design.  `key` is padded with
// +inf up to a multiple of 4.  Returns false when two keys are equal (probability ~0 for Gaussian
// deviates): the caller then falls back to a stable comparison sort.
__attribute__((target("avx2"))) static bool rank_sort_avx2(const double *key, int n, int32_t *order) {
    const int npad = (n + 3) & ~3;
    uint64_t seen[4] = {0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
        const __m256d ki = _mm256_set1_pd(key[i]);
        __m256i cnt = _mm256_setzero_si256();
        for (int j = 0; j < npad; j += 4) {
            __m256d lt = _mm256_cmp_pd(_mm256_loadu_pd(key + j), ki, _CMP_LT_OQ);
            cnt = _mm256_sub_epi64(cnt, _mm256_castpd_si256(lt));
        }
        alignas(32) int64_t c[4];
        _mm256_store_si256(reinterpret_cast<__m256i *>(c), cnt);
        const int rank = int(c[0] + c[1] + c[2] + c[3]);
        if (seen[rank >> 6] & (uint64_t(1) << (rank & 63))) return false;
        seen[rank >> 6] |= uint64_t(1) << (rank & 63);
        order[rank] = i;
    }
    return true;
}

static bool rank_sort_generic(const double *key, int n, int32_t *order) {
    uint64_t seen[4] = {0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
        const double ki = key[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += key[j] < ki;
        if (seen[rank >> 6] & (uint64_t(1) << (rank & 63))) return false;
        seen[rank >> 6] |= uint64_t(1) << (rank & 63);
        order[rank] = i;
    }
    return true;
}

// order[0..n) = argsort(key[0..n)); key must have room for 3 more entries (padding)
static void small_argsort(double *key, int n, int32_t *order, std::vector<std::pair<double, int32_t>> &scratch) {
    static const bool has_avx2 = __builtin_cpu_supports("avx2");
    if (n <= 256) {
        for (int t = n; t < ((n + 3) & ~3); ++t) key[t] = std::numeric_limits<double>::infinity();
        if (has_avx2 ? rank_sort_avx2(key, n, order) : rank_sort_generic(key, n, order)) return;
    }
    scratch.resize(size_t(n));
    for (int t = 0; t < n; ++t) scratch[size_t(t)] = {key[t], t};
    std::sort(scratch.begin(), scratch.end());  // pairs: ties broken by index
    for (int t = 0; t < n; ++t) order[t] = scratch[size_t(t)].second;
}

}  // namespace hostperm
}  // namespace cna

using namespace cna;
using namespace cna::hostperm;

extern "C" {

int cna_host_randn(uint32_t *key, int *pos, int *has_gauss, double *gauss, int64_t count, double *out,
                   int n_threads) {
    CNA_REQUIRE(key && pos && has_gauss && gauss && out && count >= 0 && *pos >= 0 && *pos <= kN,
                "cna_host_randn: bad arguments");
    if (count == 0) return CNA_OK;
    if (n_threads <= 0) n_threads = default_threads();
    const double t_begin = now_ms();
    int64_t first = 0;
    if (*has_gauss) {  // legacy_gauss: the cached deviate goes out first
        out[first++] = *gauss;
        *has_gauss = 0;
        *gauss = 0.0;
    }
    // Pair p (the p-th ACCEPTED attempt) fills out[first + 2p] = f*x2 and out[first + 2p + 1] = f*x1;
    // if the count is odd the very last f*x1 stays cached in the state.
    const int64_t n_pairs = (count - first + 1) / 2;
    if (n_pairs == 0) return CNA_OK;
    const bool odd = ((count - first) & 1) != 0;

    // The only serial part is the state recurrence.  Every attempt consumes exactly four words,
    // accepted or not, so attempt j sits at words [4j, 4j+4) of the stream that starts at
    // key[pos]: all blocks that can be needed are generated first (kept untempered: each is a
    // candidate final state), then tempering, conversion, the acceptance test, the compaction and the
    // log/sqrt transform run on all threads.
    static std::mutex ws_mutex;
    static std::vector<uint32_t> raw;      // [block -1 = the caller's key][block 0][block 1]...
    static std::vector<int64_t> chunk_acc;
    std::lock_guard<std::mutex> ws_lock(ws_mutex);
    const int pos0 = *pos;
    int64_t n_blocks = 0;                  // new blocks generated so far
    int64_t j_stop = -1;                   // index of the attempt that completes pair n_pairs-1
    int64_t n_att = 0, n_chunks = 0, per_chunk = 0;
    double t_serial = 0.0;
    for (int round = 0;; ++round) {
        // expected attempts = pairs / (pi/4); 0.4 % + 256 of slack covers > 6 sigma, else loop again
        const int64_t want_att = int64_t(double(n_pairs) * (1.2732395447351628 * (1.004 + 0.01 * round))) + 256;
        const int64_t want_words = 4 * want_att;
        const int64_t need_blocks = std::max<int64_t>(0, (want_words - (kN - pos0) + kN - 1) / kN);
        const double t0 = now_ms();
        if (raw.size() < size_t(need_blocks + 1) * kN) raw.resize(size_t(need_blocks + 1) * kN);
        if (n_blocks == 0) std::copy(key, key + kN, raw.begin());
        for (; n_blocks < need_blocks; ++n_blocks)
            mt_next_block(raw.data() + size_t(n_blocks) * kN, raw.data() + size_t(n_blocks + 1) * kN);
        t_serial += now_ms() - t0;
        const uint32_t *stream = raw.data() + pos0;   // word i of the stream
        n_att = ((kN - pos0) + n_blocks * kN) / 4;
        per_chunk = std::max<int64_t>(4096, (n_att + 8 * n_threads - 1) / (8 * n_threads));
        n_chunks = (n_att + per_chunk - 1) / per_chunk;
        chunk_acc.assign(size_t(n_chunks) + 1, 0);
        parallel_for(n_chunks, n_threads, [&](int64_t ca, int64_t cb) {  // pass 1: acceptances per chunk
            for (int64_t c = ca; c < cb; ++c) {
                const int64_t j0 = c * per_chunk, j1 = std::min(n_att, j0 + per_chunk);
                int64_t acc = 0;
                for (int64_t j = j0; j < j1; ++j) {
                    double x1, x2;
                    double r2 = polar_attempt(stream + 4 * j, x1, x2);
                    acc += !(r2 >= 1.0 || r2 == 0.0);
                }
                chunk_acc[size_t(c) + 1] = acc;
            }
        });
        for (int64_t c = 0; c < n_chunks; ++c) chunk_acc[size_t(c) + 1] += chunk_acc[size_t(c)];
        if (chunk_acc[size_t(n_chunks)] >= n_pairs) break;
    }
    const uint32_t *stream = raw.data() + pos0;
    double last_second = 0.0;
    std::vector<int64_t> stops(size_t(n_chunks), -1);
    parallel_for(n_chunks, n_threads, [&](int64_t ca, int64_t cb) {  // pass 2: compaction + transform
        for (int64_t c = ca; c < cb; ++c) {
            int64_t p = chunk_acc[size_t(c)];
            if (p >= n_pairs) break;
            const int64_t j0 = c * per_chunk, j1 = std::min(n_att, j0 + per_chunk);
            for (int64_t j = j0; j < j1 && p < n_pairs; ++j) {
                double x1, x2;
                double r2 = polar_attempt(stream + 4 * j, x1, x2);
                if (r2 >= 1.0 || r2 == 0.0) continue;
                double f = sqrt(-2.0 * log(r2) / r2);
                out[first + 2 * p] = f * x2;
                if (p == n_pairs - 1) {
                    stops[size_t(c)] = j;
                    if (odd) last_second = f * x1;
                    else out[first + 2 * p + 1] = f * x1;
                } else {
                    out[first + 2 * p + 1] = f * x1;
                }
                ++p;
            }
        }
    });
    for (int64_t c = 0; c < n_chunks; ++c)
        if (stops[size_t(c)] >= 0) j_stop = stops[size_t(c)];
    if (odd) {
        *gauss = last_second;
        *has_gauss = 1;
    }
    // state after the last consumed word (numpy regenerates lazily: pos may be left at 624)
    const int64_t consumed = 4 * (j_stop + 1);
    if (consumed <= kN - pos0) {
        *pos = pos0 + int(consumed);  // still inside the caller's block: key unchanged
    } else {
        const int64_t c = consumed - (kN - pos0);
        const int64_t blk = (c - 1) / kN;  // new block that holds the last consumed word
        std::copy(raw.begin() + size_t(blk + 1) * kN, raw.begin() + size_t(blk + 2) * kN, key);
        *pos = int(c - blk * kN);
    }
    if (timing_on())
        fprintf(stderr, "[cna timing] host_randn: %lld deviates, state recurrence %.2f ms (serial), the rest %.2f ms (%d threads)\n",
                (long long)count, t_serial, now_ms() - t_begin - t_serial, n_threads);
    return CNA_OK;
}

int cna_host_perm_blocks(uint32_t *key, int *pos, int *has_gauss, double *gauss, int n_blocks,
                         const int32_t *block_off, const int32_t *src_pos, int64_t num, int32_t *out,
                         int64_t ld_out, int n_threads) {
    CNA_REQUIRE(n_blocks >= 0 && block_off && out && num >= 0, "cna_host_perm_blocks: bad arguments");
    if (n_threads <= 0) n_threads = default_threads();
    const int64_t total_rows = block_off[n_blocks];
    CNA_REQUIRE(ld_out >= total_rows, "cna_host_perm_blocks: ld_out too small");
    if (total_rows == 0 || num == 0) return CNA_OK;
    // np.random.randn(rows_b, num) for each block, in order: one continuous stream
    const size_t nz = static_cast<size_t>(total_rows) * static_cast<size_t>(num);
    static std::mutex z_mutex;
    static std::vector<double> z;
    std::lock_guard<std::mutex> z_lock(z_mutex);
    if (z.size() < nz) z.resize(nz);
    int rc = cna_host_randn(key, pos, has_gauss, gauss, total_rows * num, z.data(), n_threads);
    if (rc != CNA_OK) return rc;
    // argsort(axis=0) per column of each block; ties have probability zero, so any comparison sort
    // gives numpy's answer.  Result rows are permutations k, columns are positions.
    // Columns are walked in tiles of 16 so that every cache line of the row-major block is read once
    // (a single column is a stride-`num` walk: one cache and TLB miss per element).
    const double t_sort = now_ms();
    constexpr int kTile = 16;
    const int64_t n_tiles = (num + kTile - 1) / kTile;
    parallel_for(n_tiles, n_threads, [&](int64_t ta, int64_t tb) {
        std::vector<std::pair<double, int32_t>> scratch;
        std::vector<double> keys;
        std::vector<int32_t> order;
        for (int64_t tile = ta; tile < tb; ++tile) {
            const int64_t k0 = tile * kTile;
            const int w = int(std::min<int64_t>(kTile, num - k0));
            for (int blk = 0; blk < n_blocks; ++blk) {
                const int32_t r0 = block_off[blk], rows = block_off[blk + 1] - r0;
                const size_t stride = size_t(rows) + 4;  // room for the +inf padding of the rank sort
                keys.resize(stride * kTile);
                order.resize(size_t(rows));
                const double *zb = z.data() + size_t(r0) * size_t(num) + k0;  // block is [rows x num]
                for (int32_t t = 0; t < rows; ++t) {
                    const double *zr = zb + size_t(t) * size_t(num);
                    for (int c = 0; c < w; ++c) keys[size_t(c) * stride + t] = zr[c];
                }
                for (int c = 0; c < w; ++c) {
                    small_argsort(keys.data() + size_t(c) * stride, rows, order.data(), scratch);
                    int32_t *o = out + (k0 + c) * ld_out;
                    if (src_pos) {  // _stats.py:14-16: bix[bi[t], k] = bi[argsort[t]]
                        for (int32_t t = 0; t < rows; ++t) o[src_pos[r0 + t]] = src_pos[r0 + order[size_t(t)]];
                    } else {        // raw order of this block
                        for (int32_t t = 0; t < rows; ++t) o[r0 + t] = order[size_t(t)];
                    }
                }
            }
        }
    });
    if (timing_on())
        fprintf(stderr, "[cna timing] host_perm_blocks: argsort + scatter %.2f ms (%d threads)\n", now_ms() - t_sort, n_threads);
    return CNA_OK;
}

// Asynchronous form: the draw runs on a library-owned thread (no Python thread, hence no waiting for
// the interpreter lock on either side); cna_host_perm_wait joins it and returns its status.
struct PermJob {
    std::thread th;
    std::atomic<int> finished{0};
    int rc = 0;
};

void *cna_host_perm_blocks_async(uint32_t *key, int *pos, int *has_gauss, double *gauss, int n_blocks,
                                 const int32_t *block_off, const int32_t *src_pos, int64_t num, int32_t *out,
                                 int64_t ld_out, int n_threads) {
    PermJob *job = new PermJob();
    job->th = std::thread([=] {
        job->rc = cna_host_perm_blocks(key, pos, has_gauss, gauss, n_blocks, block_off, src_pos, num, out, ld_out,
                                       n_threads);
        job->finished.store(1, std::memory_order_release);
    });
    return job;
}

int cna_host_perm_done(void *handle) {
    PermJob *job = static_cast<PermJob *>(handle);
    return job ? job->finished.load(std::memory_order_acquire) : 1;
}

int cna_host_perm_wait(void *handle) {
    PermJob *job = static_cast<PermJob *>(handle);
    if (!job) return cna::set_error(CNA_ERR_INVALID, "cna_host_perm_wait: null handle");
    if (job->th.joinable()) job->th.join();
    int rc = job->rc;
    delete job;
    return rc;
}

}  // extern "C"
