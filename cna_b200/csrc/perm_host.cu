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
// (no FMA contraction: this translation unit is built with -ffp-contract=off) and use the same
// libm log/sqrt the numpy extension resolves to.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

#include "common.cuh"

namespace cna {
namespace hostperm {

constexpr int kN = 624, kM = 397;
constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;

// One block of the generator: key[624] is numpy's state vector.
static void mt_regen(uint32_t *__restrict__ key) {
    int i = 0;
    for (; i < kN - kM; ++i) {
        uint32_t y = (key[i] & kUpper) | (key[i + 1] & kLower);
        key[i] = key[i + kM] ^ (y >> 1) ^ (-(y & 1u) & kMatrixA);
    }
    for (; i < kN - 1; ++i) {
        uint32_t y = (key[i] & kUpper) | (key[i + 1] & kLower);
        key[i] = key[i + (kM - kN)] ^ (y >> 1) ^ (-(y & 1u) & kMatrixA);
    }
    uint32_t y = (key[kN - 1] & kUpper) | (key[0] & kLower);
    key[kN - 1] = key[kM - 1] ^ (y >> 1) ^ (-(y & 1u) & kMatrixA);
}

static inline uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
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
    if (n_threads <= 0) n_threads = int(std::max(1u, std::min(16u, std::thread::hardware_concurrency())));
    int64_t i = 0;
    if (*has_gauss) {  // legacy_gauss: the cached deviate goes out first
        out[i++] = *gauss;
        *has_gauss = 0;
        *gauss = 0.0;
    }
    // Serial part.  Every trip of legacy_gauss's rejection loop consumes exactly four 32-bit words
    // (two doubles), accepted or not, so "attempt j" is words [4j, 4j+4) of the stream: the words of a
    // whole generator block are tempered, converted and tested in flat (vectorisable) loops, and only
    // the compaction of accepted attempts walks them in order.  Pair p fills
    // out[first + 2p] = f*x2 and out[first + 2p + 1] = f*x1.
    const int64_t first = i, n_pairs = (count - first + 1) / 2;
    const size_t np_ = static_cast<size_t>(n_pairs);
    // workspaces persist between calls (first-touch page faults of ~50 MB of fresh vectors cost as
    // much as the generator itself); calls are serialised by the global generator state anyway
    static std::mutex ws_mutex;
    static std::vector<double> x1s, x2s, r2s;
    std::lock_guard<std::mutex> ws_lock(ws_mutex);
    if (x1s.size() < np_) {
        x1s.resize(np_);
        x2s.resize(np_);
        r2s.resize(np_);
    }
    {
        uint32_t w[kN + 4];
        double bx1[kN / 4 + 1], bx2[kN / 4 + 1], br2[kN / 4 + 1];
        int p0 = *pos, left = 0;  // `left` words carried over from the previous block sit at w[0..left)
        int64_t acc = 0;
        while (acc < n_pairs) {
            if (p0 == kN) {
                mt_regen(key);
                p0 = 0;
            }
            const int fresh = kN - p0;
            for (int t = 0; t < fresh; ++t) w[left + t] = mt_temper(key[p0 + t]);
            const int have = left + fresh, n_att = have / 4;
            for (int j = 0; j < n_att; ++j) {  // mt19937_next_double twice, then the polar test
                int32_t a1 = int32_t(w[4 * j] >> 5), b1 = int32_t(w[4 * j + 1] >> 6);
                int32_t a2 = int32_t(w[4 * j + 2] >> 5), b2 = int32_t(w[4 * j + 3] >> 6);
                double d1 = (a1 * 67108864.0 + b1) / 9007199254740992.0;
                double d2 = (a2 * 67108864.0 + b2) / 9007199254740992.0;
                double x1 = 2.0 * d1 - 1.0, x2 = 2.0 * d2 - 1.0;
                bx1[j] = x1;
                bx2[j] = x2;
                br2[j] = x1 * x1 + x2 * x2;
            }
            int j = 0;
            for (; j < n_att && acc < n_pairs; ++j) {
                double r2 = br2[j];
                if (r2 >= 1.0 || r2 == 0.0) continue;
                x1s[size_t(acc)] = bx1[j];
                x2s[size_t(acc)] = bx2[j];
                r2s[size_t(acc)] = r2;
                ++acc;
            }
            if (acc == n_pairs) {  // stopped inside this block: j attempts of it were consumed
                *pos = p0 + (4 * j - left);
                break;
            }
            left = have - 4 * n_att;
            for (int t = 0; t < left; ++t) w[t] = w[4 * n_att + t];
            p0 = kN;
        }
    }
    const bool odd = ((count - first) & 1) != 0;  // the last pair's second deviate stays cached
    parallel_for(n_pairs, n_threads, [&](int64_t a, int64_t b) {
        for (int64_t p = a; p < b; ++p) {
            double r2 = r2s[size_t(p)];
            double f = sqrt(-2.0 * log(r2) / r2);
            out[first + 2 * p] = f * x2s[size_t(p)];
            double second = f * x1s[size_t(p)];
            if (p == n_pairs - 1 && odd) {
                *gauss = second;
                *has_gauss = 1;
            } else {
                out[first + 2 * p + 1] = second;
            }
        }
    });
    return CNA_OK;
}

int cna_host_perm_blocks(uint32_t *key, int *pos, int *has_gauss, double *gauss, int n_blocks,
                         const int32_t *block_off, const int32_t *src_pos, int64_t num, int32_t *out,
                         int64_t ld_out, int n_threads) {
    CNA_REQUIRE(n_blocks >= 0 && block_off && out && num >= 0, "cna_host_perm_blocks: bad arguments");
    if (n_threads <= 0) n_threads = int(std::max(1u, std::min(16u, std::thread::hardware_concurrency())));
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
    constexpr int kTile = 16;
    const int64_t n_tiles = (num + kTile - 1) / kTile;
    parallel_for(n_tiles, n_threads, [&](int64_t ta, int64_t tb) {
        std::vector<std::pair<double, int32_t>> buf;
        for (int64_t tile = ta; tile < tb; ++tile) {
            const int64_t k0 = tile * kTile;
            const int w = int(std::min<int64_t>(kTile, num - k0));
            for (int blk = 0; blk < n_blocks; ++blk) {
                const int32_t r0 = block_off[blk], rows = block_off[blk + 1] - r0;
                buf.resize(size_t(rows) * kTile);
                const double *zb = z.data() + size_t(r0) * size_t(num) + k0;  // block is [rows x num]
                for (int32_t t = 0; t < rows; ++t) {
                    const double *zr = zb + size_t(t) * size_t(num);
                    for (int c = 0; c < w; ++c) buf[size_t(c) * rows + t] = {zr[c], t};
                }
                for (int c = 0; c < w; ++c) {
                    auto *col = buf.data() + size_t(c) * rows;
                    std::sort(col, col + rows);
                    int32_t *o = out + (k0 + c) * ld_out;
                    if (src_pos) {  // _stats.py:14-16: bix[bi[t], k] = bi[argsort[t]]
                        for (int32_t t = 0; t < rows; ++t) o[src_pos[r0 + t]] = src_pos[r0 + col[t].second];
                    } else {        // raw order of this block
                        for (int32_t t = 0; t < rows; ++t) o[r0 + t] = col[t].second;
                    }
                }
            }
        }
    });
    return CNA_OK;
}

}  // extern "C"
