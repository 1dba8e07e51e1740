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
#include <condition_variable>
#include <cstdlib>
#include <functional>
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

// A small persistent pool: the workers are created once per process and parked on a condition variable
// (spawning 15 threads costs more than a tenth of the whole draw).  run(n, fn) executes fn(worker) for
// worker = 0 .. n-1, worker 0 on the calling thread, and returns when all are done.
class Pool {
  public:
    static Pool &get() {
        // never destroyed: the parked workers outlive main(), and destroying a condition variable with
        // waiters blocks process exit
        static Pool *p = new Pool();
        return *p;
    }
    template <typename F>
    void run(int n, F fn) {
        if (n <= 1) {
            fn(0);
            return;
        }
        std::lock_guard<std::mutex> serial(run_mutex_);
        grow(n - 1);
        std::function<void(int)> job = fn;
        {
            std::lock_guard<std::mutex> lk(m_);
            job_ = &job;
            active_ = n - 1;
            remaining_ = n - 1;
            ++generation_;
        }
        cv_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [&] { return remaining_ == 0; });
        job_ = nullptr;
    }

  private:
    void grow(int n) {
        while (int(threads_.size()) < n) {
            const int id = int(threads_.size());
            threads_.emplace_back([this, id] {
                uint64_t seen = 0;
                for (;;) {
                    std::function<void(int)> *job;
                    {
                        std::unique_lock<std::mutex> lk(m_);
                        cv_.wait(lk, [&] { return generation_ != seen; });
                        seen = generation_;
                        if (id >= active_) continue;
                        job = job_;
                    }
                    (*job)(id + 1);
                    std::lock_guard<std::mutex> lk(m_);
                    if (--remaining_ == 0) done_.notify_one();
                }
            });
            threads_.back().detach();
        }
    }
    std::mutex m_, run_mutex_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> threads_;
    std::function<void(int)> *job_ = nullptr;
    uint64_t generation_ = 0;
    int active_ = 0, remaining_ = 0;
};

static inline void cpu_relax() { _mm_pause(); }

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

namespace cna {
namespace hostperm {

// Work that follows the deviates: task t may start once the first `need[t]` deviates are in place
// (the argsort of one batch block needs exactly the deviates of that block).
struct PostTasks {
    int64_t count = 0;
    std::function<int64_t(int64_t)> need;         // deviates task t depends on
    std::function<void(int64_t, int)> run;        // (task, worker)
};

// The draw as ONE pass over the stream, pipelined behind the state recurrence:
//   * the calling thread generates the MT19937 blocks (the only inherently serial part) and publishes
//     how many are ready;
//   * the workers claim chunks of attempts in stream order.  For its chunk a worker tempers, converts and
//     tests every attempt, keeps the accepted ones (r2, x1, x2) in a thread-local buffer, waits for the
//     chunk before it to publish its cumulative number of accepted pairs (a short wait: chunks are claimed
//     in order and publish before their expensive part), publishes its own, and then runs the log / sqrt
//     transform into the output positions that number fixes;
//   * when the chunks are exhausted the workers go on to the post tasks (per-column argsorts), each of
//     which waits only for the deviates it reads.
// Every attempt consumes exactly four words, accepted or not, so attempt j sits at words [4j, 4j + 4)
// of the stream that starts at key[pos]; accepted attempt number p fills out[first + 2p] = f*x2 and
// out[first + 2p + 1] = f*x1 (legacy_gauss returns the cached second deviate on the next call).
static int randn_pipeline(uint32_t *key, int *pos, int *has_gauss, double *gauss, int64_t count, double *out,
                          int n_threads, const PostTasks *post) {
    const double t_begin = now_ms();
    int64_t first = 0;
    if (count > 0 && *has_gauss) {  // legacy_gauss: the cached deviate goes out first
        out[first++] = *gauss;
        *has_gauss = 0;
        *gauss = 0.0;
    }
    const int64_t n_pairs = count > first ? (count - first + 1) / 2 : 0;
    const bool odd = ((count - first) & 1) != 0;
    const int pos0 = *pos;
    static std::mutex ws_mutex;
    static std::vector<uint32_t> raw;  // [the caller's key][block 0][block 1]...: untempered, each a candidate final state
    std::lock_guard<std::mutex> ws_lock(ws_mutex);

    constexpr int64_t kChunk = 2048;  // attempts per chunk (8192 words, ~13 blocks)
    struct Accepted {
        double r2, x1, x2;
        int32_t j;  // attempt index inside the chunk
    };
    std::atomic<int64_t> blocks_ready{0}, next_chunk{0}, next_task{0}, deviates_done{0}, j_stop{-1};
    std::atomic<int> failed{0};
    double last_second = 0.0;
    int64_t n_blocks = 0, chunks_total = 0;
    std::vector<std::atomic<int64_t>> cum;   // cum[c] = accepted pairs before chunk c (-1 = not yet known)
    std::vector<std::atomic<int8_t>> chunk_done;
    double t_serial = 0.0;

    auto words_available = [&](int64_t blocks) { return int64_t(kN - pos0) + blocks * kN; };

    for (int round = 0; n_pairs > 0; ++round) {
        // expected attempts = pairs / (pi/4); 0.4 % + 256 of slack covers > 6 sigma, else one more round
        const int64_t want_att = int64_t(double(n_pairs) * (1.2732395447351628 * (1.004 + 0.01 * round))) + 256;
        const int64_t need_blocks = std::max<int64_t>(0, (4 * want_att - (kN - pos0) + kN - 1) / kN);
        if (raw.size() < size_t(need_blocks + 1) * kN) {
            std::vector<uint32_t> bigger(size_t(need_blocks + 1) * kN);
            std::copy(raw.begin(), raw.begin() + std::min(raw.size(), size_t(n_blocks + 1) * kN), bigger.begin());
            raw.swap(bigger);
        }
        if (n_blocks == 0) std::copy(key, key + kN, raw.begin());
        const int64_t n_att = words_available(need_blocks) / 4;
        const int64_t n_chunks = (n_att + kChunk - 1) / kChunk;
        // a later round re-runs everything over the longer stream (it practically never happens)
        cum = std::vector<std::atomic<int64_t>>(size_t(n_chunks) + 1);
        chunk_done = std::vector<std::atomic<int8_t>>(size_t(n_chunks));
        for (auto &c : cum) c.store(-1, std::memory_order_relaxed);
        for (auto &c : chunk_done) c.store(0, std::memory_order_relaxed);
        cum[0].store(0, std::memory_order_relaxed);
        next_chunk.store(0);
        blocks_ready.store(n_blocks);
        deviates_done.store(first);
        chunks_total = n_chunks;
        const uint32_t *stream = raw.data() + pos0;

        auto recurrence = [&]() {
            const double t0 = now_ms();
            for (; n_blocks < need_blocks; ++n_blocks) {
                mt_next_block(raw.data() + size_t(n_blocks) * kN, raw.data() + size_t(n_blocks + 1) * kN);
                if ((n_blocks & 7) == 7) blocks_ready.store(n_blocks + 1, std::memory_order_release);
            }
            blocks_ready.store(n_blocks, std::memory_order_release);
            t_serial += now_ms() - t0;
        };
        auto chunk_loop = [&]() {
            std::vector<Accepted> acc(static_cast<size_t>(kChunk));
            for (;;) {
                const int64_t c = next_chunk.fetch_add(1, std::memory_order_relaxed);
                if (c >= n_chunks) break;
                const int64_t j0 = c * kChunk, j1 = std::min(n_att, j0 + kChunk);
                while (words_available(blocks_ready.load(std::memory_order_acquire)) < 4 * j1) cpu_relax();
                int64_t na = 0;
                for (int64_t j = j0; j < j1; ++j) {
                    double x1, x2;
                    const double r2 = polar_attempt(stream + 4 * j, x1, x2);
                    if (r2 >= 1.0 || r2 == 0.0) continue;
                    acc[size_t(na++)] = {r2, x1, x2, int32_t(j - j0)};
                }
                int64_t p0;
                while ((p0 = cum[size_t(c)].load(std::memory_order_acquire)) < 0) cpu_relax();
                cum[size_t(c) + 1].store(p0 + na, std::memory_order_release);
                for (int64_t i = 0; i < na && p0 + i < n_pairs; ++i) {
                    const Accepted &e = acc[size_t(i)];
                    const int64_t p = p0 + i;
                    const double f = sqrt(-2.0 * log(e.r2) / e.r2);
                    out[first + 2 * p] = f * e.x2;
                    if (p == n_pairs - 1) {
                        j_stop.store(j0 + e.j, std::memory_order_relaxed);
                        if (odd) last_second = f * e.x1;
                        else out[first + 2 * p + 1] = f * e.x1;
                    } else {
                        out[first + 2 * p + 1] = f * e.x1;
                    }
                }
                chunk_done[size_t(c)].store(1, std::memory_order_release);
            }
        };
        // The post tasks run inside the same parallel region, after the chunks.  Task t waits until every
        // chunk that holds one of its deviates is done; if the round turns out to be short of attempts
        // (practically never) the tasks are abandoned and the next round starts over.
        auto task_loop = [&](int w) {
            if (!post || post->count == 0) return;
            int64_t mark = 0;  // chunks [0, mark) are known to be done
            for (;;) {
                const int64_t t = next_task.fetch_add(1, std::memory_order_relaxed);
                if (t >= post->count) break;
                const int64_t need = std::min(post->need(t), count);  // deviates [0, need) must be in place
                for (;;) {
                    if (failed.load(std::memory_order_relaxed)) return;
                    while (mark < n_chunks && chunk_done[size_t(mark)].load(std::memory_order_acquire)) ++mark;
                    if (mark >= n_chunks) {
                        if (cum[size_t(n_chunks)].load(std::memory_order_acquire) < n_pairs) {
                            failed.store(1);
                            return;
                        }
                        break;
                    }
                    // chunks [0, mark) done: pairs [0, cum[mark]) are in place
                    const int64_t pairs_done = std::min(cum[size_t(mark)].load(std::memory_order_acquire), n_pairs);
                    if (first + 2 * pairs_done >= need) break;
                    cpu_relax();
                }
                post->run(t, w);
            }
        };
        next_task.store(0);
        failed.store(0);
        Pool::get().run(n_threads, [&](int w) {
            if (w == 0) recurrence();
            chunk_loop();
            task_loop(w);
        });
        if (cum[size_t(n_chunks)].load() >= n_pairs) break;
    }
    if (n_pairs > 0) {
        if (odd) {
            *gauss = last_second;
            *has_gauss = 1;
        }
        // state after the last consumed word (numpy regenerates lazily: pos may be left at 624)
        const int64_t consumed = 4 * (j_stop.load() + 1);
        if (consumed <= kN - pos0) {
            *pos = pos0 + int(consumed);  // still inside the caller's block: key unchanged
        } else {
            const int64_t c = consumed - (kN - pos0);
            const int64_t blk = (c - 1) / kN;  // new block that holds the last consumed word
            std::copy(raw.begin() + size_t(blk + 1) * kN, raw.begin() + size_t(blk + 2) * kN, key);
            *pos = int(c - blk * kN);
        }
    } else if (post && post->count > 0) {  // every deviate came out of the cache: nothing to wait for
        next_task.store(0);
        Pool::get().run(n_threads, [&](int w) {
            for (;;) {
                const int64_t t = next_task.fetch_add(1, std::memory_order_relaxed);
                if (t >= post->count) break;
                post->run(t, w);
            }
        });
    }
    (void)chunks_total;
    if (timing_on())
        fprintf(stderr, "[cna timing] host draw: %lld deviates%s, state recurrence %.2f ms (overlapped), total %.2f ms (%d threads)\n",
                (long long)count, post ? " + argsorts" : "", t_serial, now_ms() - t_begin, n_threads);
    return CNA_OK;
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
    return randn_pipeline(key, pos, has_gauss, gauss, count, out, n_threads, nullptr);
}

int cna_host_perm_blocks(uint32_t *key, int *pos, int *has_gauss, double *gauss, int n_blocks,
                         const int32_t *block_off, const int32_t *src_pos, int64_t num, int32_t *out,
                         int64_t ld_out, int n_threads) {
    CNA_REQUIRE(n_blocks >= 0 && block_off && out && num >= 0 && key && pos && has_gauss && gauss,
                "cna_host_perm_blocks: bad arguments");
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
    // argsort(axis=0) per column of each block; ties have probability zero, so any comparison sort
    // gives numpy's answer.  Result rows are permutations k, columns are positions.
    // Columns are walked in tiles of 16 so that every cache line of the row-major block is read once
    // (a single column is a stride-`num` walk: one cache and TLB miss per element).  Task = (block,
    // tile), blocks first: the argsorts of a batch start as soon as its deviates are in place.
    constexpr int kTile = 16;
    const int64_t n_tiles = (num + kTile - 1) / kTile;
    struct Scratch {
        std::vector<std::pair<double, int32_t>> pairs;
        std::vector<double> keys;
        std::vector<int32_t> order;
    };
    std::vector<Scratch> scratch(size_t(n_threads) + 1);
    PostTasks post;
    post.count = int64_t(n_blocks) * n_tiles;
    post.need = [&](int64_t t) { return int64_t(block_off[t / n_tiles + 1]) * num; };
    post.run = [&](int64_t t, int w) {
        Scratch &sc = scratch[size_t(w)];
        const int blk = int(t / n_tiles);
        const int64_t k0 = (t % n_tiles) * kTile;
        const int wdt = int(std::min<int64_t>(kTile, num - k0));
        const int32_t r0 = block_off[blk], rows = block_off[blk + 1] - r0;
        if (rows == 0) return;
        const size_t stride = size_t(rows) + 4;  // room for the +inf padding of the rank sort
        sc.keys.resize(stride * kTile);
        sc.order.resize(size_t(rows));
        const double *zb = z.data() + size_t(r0) * size_t(num) + k0;  // block is [rows x num]
        for (int32_t tt = 0; tt < rows; ++tt) {
            const double *zr = zb + size_t(tt) * size_t(num);
            for (int c = 0; c < wdt; ++c) sc.keys[size_t(c) * stride + tt] = zr[c];
        }
        for (int c = 0; c < wdt; ++c) {
            small_argsort(sc.keys.data() + size_t(c) * stride, rows, sc.order.data(), sc.pairs);
            int32_t *o = out + (k0 + c) * ld_out;
            if (src_pos) {  // _stats.py:14-16: bix[bi[t], k] = bi[argsort[t]]
                for (int32_t tt = 0; tt < rows; ++tt) o[src_pos[r0 + tt]] = src_pos[r0 + sc.order[size_t(tt)]];
            } else {        // raw order of this block
                for (int32_t tt = 0; tt < rows; ++tt) o[r0 + tt] = sc.order[size_t(tt)];
            }
        }
    };
    return randn_pipeline(key, pos, has_gauss, gauss, total_rows * num, z.data(), n_threads, &post);
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
