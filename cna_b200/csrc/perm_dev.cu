// The permutation draw on the device: numpy's legacy generator (MT19937 + polar Gaussian) restated for
// the GPU, followed by the per-column argsort of the reference.
//
// Reference: src/cna/tools/_stats.py:11-16 (conditional_permutation) and :31 (grouplevel_permutation):
//     np.argsort(np.random.randn(rows_b, num), axis=0)   for every batch block b.
// The host restatement (perm_host.cpp) is bit-exact but needs all host threads for ~4 ms per 10 000 x 200
// draw, and on a box where eight ranks share the host it becomes the critical path of a call.  Here:
//   1. mt_stream_kernel (one CTA): the MT19937 state recurrence, the only serial part — x[k + 624]
//      depends on x[k], x[k + 1], x[k + 397], so 227 consecutive words are produced per wave; the
//      untempered stream is written to global memory block by block (every block is a candidate final
//      state, exactly like numpy's key array);
//   2. attempts_count_kernel / attempts_emit_kernel: attempt j of legacy_gauss's rejection loop consumes
//      words [4 j, 4 j + 4) of the stream whether it is accepted or not, so every attempt is tested in
//      parallel, the accepted ones are numbered by a scan, and accepted pair p yields deviates
//      first + 2 p (= f x2) and first + 2 p + 1 (= f x1), f = sqrt(-2 log(r2) / r2);
//   3. rank_kernel (a warp per permutation): argsort of every batch block by counting, and the scatter
//      through the sample positions of the block (_stats.py:14-16).
// What is bit-exact and what is guarded: the integer stream, the acceptance tests and the uniform
// doubles are exact (integer and correctly rounded IEEE arithmetic).  CUDA's log() may differ from the
// host libm in the last place (both are accurate to < 1 ulp), which moves a deviate by at most ~3 ulp —
// it cannot change an argsort unless two keys of a column are closer than that.  rank_kernel therefore
// checks every pair of neighbours in sorted order and raises `ambiguous` when a gap is below 1e-14
// relative (probability ~1e-6 per call); the caller then repeats the draw with the host engine.  A cached
// second deviate (odd count) is finished by the caller on the host from (r2, x1), so the generator state
// left behind is bit-exact as well.
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace cna {
namespace permdev {

constexpr int kN = 624, kM = 397;
constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;

__device__ __forceinline__ uint32_t twist(uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t y = (a & kUpper) | (b & kLower);
    return c ^ (y >> 1) ^ ((y & 1u) ? kMatrixA : 0u);
}
__device__ __forceinline__ uint32_t temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// raw[0 .. 624) = the caller's key; raw[624 b ..] = block b (b >= 1).  256 threads.
__global__ void __launch_bounds__(256) mt_stream_kernel(uint32_t *__restrict__ raw, int64_t n_blocks) {
    __shared__ uint32_t st[2][kN + 8];
    const int t = threadIdx.x;
    for (int i = t; i < kN; i += 256) st[0][i] = raw[i];
    __syncthreads();
    int cur = 0;
    for (int64_t b = 1; b <= n_blocks; ++b) {
        const uint32_t *o = st[cur];
        uint32_t *nw = st[cur ^ 1];
        uint32_t *out = raw + b * kN;  // every word goes to global memory as it is produced
        if (t < kN - kM) out[t] = nw[t] = twist(o[t], o[t + 1], o[t + kM]);  // [0, 227): old words only
        __syncthreads();
        {
            const int i = (kN - kM) + t;  // [227, 454): new words written one wave earlier
            if (t < kN - kM) out[i] = nw[i] = twist(o[i], o[i + 1], nw[i - (kN - kM)]);
        }
        __syncthreads();
        {
            const int i = 2 * (kN - kM) + t;  // [454, 623), and the wrap-around word 623
            if (i < kN - 1) out[i] = nw[i] = twist(o[i], o[i + 1], nw[i - (kN - kM)]);
            else if (i == kN - 1) out[i] = nw[i] = twist(o[kN - 1], nw[0], nw[kM - 1]);
        }
        __syncthreads();
        cur ^= 1;
    }
}

struct Attempt {
    double r2, x1, x2;
    bool ok;
};
// attempt j of the rejection loop = untempered words w[0..3] (numpy: mt19937_next_double twice)
__device__ __forceinline__ Attempt attempt(const uint32_t *__restrict__ w) {
    const int32_t a1 = int32_t(temper(w[0]) >> 5), b1 = int32_t(temper(w[1]) >> 6);
    const int32_t a2 = int32_t(temper(w[2]) >> 5), b2 = int32_t(temper(w[3]) >> 6);
    const double d1 = (a1 * 67108864.0 + b1) / 9007199254740992.0;
    const double d2 = (a2 * 67108864.0 + b2) / 9007199254740992.0;
    Attempt r;
    r.x1 = 2.0 * d1 - 1.0;
    r.x2 = 2.0 * d2 - 1.0;
    r.r2 = __dadd_rn(__dmul_rn(r.x1, r.x1), __dmul_rn(r.x2, r.x2));  // no FMA contraction: numpy rounds both products
    r.ok = !(r.r2 >= 1.0 || r.r2 == 0.0);
    return r;
}

constexpr int kChunk = 1024;  // attempts per CTA (256 threads x 4)

__global__ void __launch_bounds__(256) attempts_count_kernel(const uint32_t *__restrict__ stream, int64_t n_att,
                                                             int32_t *__restrict__ chunk_count) {
    const int64_t j0 = int64_t(blockIdx.x) * kChunk + threadIdx.x * 4;
    int c = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (j0 + u < n_att) c += attempt(stream + 4 * (j0 + u)).ok ? 1 : 0;
    __shared__ int ws[8];
    c = __reduce_add_sync(kFull, c);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) chunk_count[blockIdx.x] = ws[0] + ws[1] + ws[2] + ws[3] + ws[4] + ws[5] + ws[6] + ws[7];
}

// exclusive scan of the chunk counts (one CTA), in place as int64; total -> chunk_off[n_chunks]
__global__ void __launch_bounds__(1024) chunk_scan_kernel(const int32_t *__restrict__ chunk_count, int64_t n_chunks,
                                                          int64_t *__restrict__ chunk_off) {
    __shared__ int64_t warp_tot[32];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int64_t base = 0; base < n_chunks; base += 1024) {
        const int64_t i = base + threadIdx.x;
        int64_t v = i < n_chunks ? chunk_count[i] : 0, incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t up = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) warp_tot[w] = incl;
        __syncthreads();
        int64_t before = carry;
        for (int q = 0; q < w; ++q) before += warp_tot[q];
        if (i < n_chunks) chunk_off[i] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) chunk_off[n_chunks] = carry;
}

struct Emit {
    const uint32_t *stream;   // untempered words, attempt j at [4 j, 4 j + 4)
    int64_t n_att;
    const int64_t *chunk_off; // accepted pairs before each chunk
    int64_t first;            // deviates that precede the pairs (0 or 1: the caller's cached deviate)
    int64_t count;            // deviates wanted
    int64_t n_pairs;          // pairs needed = ceil((count - first) / 2)
    double *dev;              // [num x n] deviates, transposed: deviate of (block row t, permutation k) at k n + t
    const int64_t *blk_dev0;  // [n_blocks + 1] first deviate index of every batch block (block_off[b] * num)
    const int32_t *block_off; // [n_blocks + 1]
    int n_blocks;
    int64_t num, n;
    double *tail;             // [4]: attempt index of the last pair, its r2 and x1, unused
};

__device__ __forceinline__ void put_deviate(const Emit &e, int64_t d, double value) {
    if (d >= e.count) return;  // the second deviate of the last pair when the count is odd: cached by the caller
    int b = 0;
    while (b + 1 < e.n_blocks && d >= e.blk_dev0[b + 1]) ++b;
    const int64_t local = d - e.blk_dev0[b];
    const int64_t r = local / e.num, k = local - r * e.num;  // randn(rows_b, num) is row-major
    e.dev[k * e.n + e.block_off[b] + r] = value;
}

__global__ void __launch_bounds__(256) attempts_emit_kernel(Emit e) {
    const int64_t j0 = int64_t(blockIdx.x) * kChunk + threadIdx.x * 4;
    Attempt at[4];
    int c = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        at[u].ok = false;
        if (j0 + u < e.n_att) at[u] = attempt(e.stream + 4 * (j0 + u));
        c += at[u].ok ? 1 : 0;
    }
    // exclusive scan of c over the CTA
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += up;
    }
    __shared__ int ws[8];
    if (lane == 31) ws[w] = incl;
    __syncthreads();
    int before = 0;
    for (int q = 0; q < w; ++q) before += ws[q];
    int64_t p = e.chunk_off[blockIdx.x] + before + incl - c;  // pair index of this thread's first accepted attempt
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        if (!at[u].ok) continue;
        if (p < e.n_pairs) {
            const double f = sqrt(-2.0 * log(at[u].r2) / at[u].r2);  // legacy_gauss
            put_deviate(e, e.first + 2 * p, f * at[u].x2);
            put_deviate(e, e.first + 2 * p + 1, f * at[u].x1);
            if (p == e.n_pairs - 1) {
                e.tail[0] = double(j0 + u);
                e.tail[1] = at[u].r2;
                e.tail[2] = at[u].x1;
            }
        }
        ++p;
    }
}

// The generator state after the draw: key block and position (numpy regenerates lazily, so pos may be 624).
// state_out: [624] key words, then pos, then the number of accepted pairs found (for the caller's check).
__global__ void final_state_kernel(const uint32_t *__restrict__ raw, int pos0, const double *__restrict__ tail,
                                   const int64_t *__restrict__ chunk_off, int64_t n_chunks, int64_t n_pairs,
                                   uint32_t *__restrict__ state_out) {
    const int64_t have = chunk_off[n_chunks];
    int64_t P = pos0;  // absolute word position (from the start of the caller's key) of the next unread word
    if (n_pairs > 0) P += 4 * (int64_t(tail[0]) + 1);
    const int64_t blk = P <= kN ? 0 : (P - 1) / kN;
    for (int i = threadIdx.x; i < kN; i += blockDim.x) state_out[i] = raw[blk * kN + i];
    if (threadIdx.x == 0) {
        state_out[kN] = uint32_t(P - blk * kN);
        state_out[kN + 1] = have >= n_pairs ? 1u : 0u;  // 0: the stream was too short (practically never)
    }
}

// argsort of every batch block of every permutation by counting, then the scatter of _stats.py:14-16.
// dev: [num x n] keys (row k = permutation k); out: [num x ld_out].  A warp per permutation.
__global__ void __launch_bounds__(256) rank_kernel(const double *__restrict__ dev, int64_t num, int n, int n_blocks,
                                                   const int32_t *__restrict__ block_off,
                                                   const int32_t *__restrict__ src_pos, int32_t *__restrict__ out,
                                                   int64_t ld_out, int *__restrict__ ambiguous) {
    extern __shared__ double keys_all[];  // [8 warps][n] keys, then [8][n] sorted keys
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double *key = keys_all + size_t(w) * n, *sorted = keys_all + size_t(8 + w) * n;
    for (int64_t k = int64_t(blockIdx.x) * 8 + w; k < num; k += int64_t(gridDim.x) * 8) {
        for (int t = lane; t < n; t += 32) key[t] = dev[k * n + t];
        __syncwarp();
        for (int b = 0; b < n_blocks; ++b) {
            const int r0 = block_off[b], rows = block_off[b + 1] - r0;
            for (int t = lane; t < rows; t += 32) {
                const double kt = key[r0 + t];
                int rank = 0;
                for (int u = 0; u < rows; ++u) {
                    const double ku = key[r0 + u];
                    rank += (ku < kt || (ku == kt && u < t)) ? 1 : 0;
                }
                sorted[r0 + rank] = kt;
                // out[k, src_pos[r0 + rank]] = src_pos[r0 + t]  (position `rank` of the sorted block holds row t)
                const int dst = src_pos ? src_pos[r0 + rank] : r0 + rank;
                out[k * ld_out + dst] = src_pos ? src_pos[r0 + t] : t;
            }
            __syncwarp();
            for (int t = lane; t + 1 < rows; t += 32) {  // neighbours in sorted order: is the order beyond doubt?
                const double a = sorted[r0 + t], c = sorted[r0 + t + 1];
                if (!(c - a > 1e-14 * (fabs(a) + fabs(c)))) atomicOr(ambiguous, 1);
            }
            __syncwarp();
        }
    }
}

}  // namespace permdev
}  // namespace cna

using namespace cna;
using namespace cna::permdev;

// Workspace layout (bytes): raw stream | chunk counts | chunk offsets | deviates | tables
static int64_t attempts_for(int64_t n_pairs) {
    // expected attempts = pairs / (pi / 4); 0.4 % + 256 of slack covers > 6 sigma
    return int64_t(double(n_pairs) * (1.2732395447351628 * 1.004)) + 256;
}
static inline int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

extern "C" int64_t cna_perm_draw_workspace(int64_t n, int64_t num, int n_blocks) {
    if (n < 1 || num < 1 || n_blocks < 1) return 0;
    const int64_t count = n * num, n_pairs = (count + 1) / 2, n_att = attempts_for(n_pairs);
    const int64_t blocks = (4 * n_att + kN - 1) / kN + 2, n_chunks = (n_att + kChunk - 1) / kChunk;
    int64_t bytes = align256(4 * (blocks + 1) * kN);   // raw
    bytes = align256(bytes + 4 * n_chunks);             // chunk counts
    bytes = align256(bytes + 8 * (n_chunks + 1));       // chunk offsets
    bytes = align256(bytes + 8 * count);                // deviates
    bytes = align256(bytes + 8 * (n_blocks + 1) + 4 * (n_blocks + 1) + 4 * n + 64);  // tables
    return bytes + 1024;
}

// A page-locked staging buffer owned by the library: the caller's (pageable) state and tables are packed
// into it and travel in ONE truly asynchronous copy — a pageable cudaMemcpyAsync would wait for everything
// queued on the stream before it.
namespace {
struct Staging {
    char *host = nullptr;
    size_t bytes = 0;
    cudaEvent_t copied = nullptr;
    std::mutex m;
};
Staging g_stage;
}  // namespace

// key_host/pos/has_gauss/gauss: numpy's legacy state BEFORE the draw.  block_off_host [n_blocks + 1] and
// src_pos_host [n] (or NULL): host int32 tables as in cna_host_perm_blocks.
// state_out (device, 626 uint32): key after the draw, pos, flag "enough attempts"; tail (device, 4 doubles):
// [attempt index, r2, x1] of the last accepted pair (the caller finishes a cached deviate on the host);
// ambiguous (device int): set when an argsort could depend on the last place of log().
extern "C" int cna_perm_draw_device(const uint32_t *key_host, int pos, int has_gauss, double gauss, int n_blocks,
                                    const int32_t *block_off_host, const int32_t *src_pos_host, int64_t num,
                                    int32_t *out, int64_t ld_out, uint32_t *state_out, double *tail, int *ambiguous,
                                    void *workspace, int64_t workspace_bytes, void *stream) {
    CNA_REQUIRE(key_host && block_off_host && out && state_out && tail && ambiguous && workspace,
                "cna_perm_draw_device: null argument");
    CNA_REQUIRE(n_blocks >= 1 && n_blocks <= 64 && num >= 1 && pos >= 0 && pos <= kN,
                "cna_perm_draw_device: bad arguments (1..64 batch blocks)");
    const int64_t n = block_off_host[n_blocks];
    CNA_REQUIRE(n >= 1 && n <= 1600 && ld_out >= n, "cna_perm_draw_device: 1..1600 samples (got %lld)", (long long)n);
    CNA_REQUIRE(workspace_bytes >= cna_perm_draw_workspace(n, num, n_blocks), "cna_perm_draw_device: workspace too small");
    cudaStream_t st = as_stream(stream);
    const int64_t count = n * num, first = has_gauss ? 1 : 0;
    const int64_t n_pairs = (count - first + 1) / 2, n_att = attempts_for(n_pairs);
    const int64_t need_words = 4 * n_att - (kN - pos);
    const int64_t blocks = need_words > 0 ? (need_words + kN - 1) / kN : 0, n_chunks = (n_att + kChunk - 1) / kChunk;
    char *base = static_cast<char *>(workspace);
    uint32_t *raw = reinterpret_cast<uint32_t *>(base);
    int64_t off = align256(4 * (blocks + 1) * kN);
    int32_t *chunk_count = reinterpret_cast<int32_t *>(base + off);
    off = align256(off + 4 * n_chunks);
    int64_t *chunk_off = reinterpret_cast<int64_t *>(base + off);
    off = align256(off + 8 * (n_chunks + 1));
    double *dev = reinterpret_cast<double *>(base + off);
    off = align256(off + 8 * count);
    char *tables = base + off;  // [blk_dev0: (n_blocks + 1) int64][gauss: double][block_off][src_pos]
    const size_t t_dev0 = 0, t_gauss = 8 * size_t(n_blocks + 1), t_off = t_gauss + 8,
                 t_pos = t_off + 4 * size_t(n_blocks + 1), t_key = (t_pos + 4 * size_t(n) + 15) / 16 * 16,
                 t_bytes = t_key;  // the key goes straight to raw[0 .. 624)
    {
        std::lock_guard<std::mutex> lk(g_stage.m);
        const size_t need = t_bytes + sizeof(uint32_t) * kN;
        if (g_stage.copied) CNA_CUDA(cudaEventSynchronize(g_stage.copied));  // the previous draw's copy has left
        if (g_stage.bytes < need) {
            if (g_stage.host) cudaFreeHost(g_stage.host);
            CNA_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&g_stage.host), need, cudaHostAllocDefault));
            g_stage.bytes = need;
        }
        if (!g_stage.copied) CNA_CUDA(cudaEventCreateWithFlags(&g_stage.copied, cudaEventDisableTiming));
        char *h = g_stage.host;
        int64_t *h_dev0 = reinterpret_cast<int64_t *>(h + t_dev0);
        for (int b = 0; b <= n_blocks; ++b) h_dev0[b] = int64_t(block_off_host[b]) * num;
        *reinterpret_cast<double *>(h + t_gauss) = gauss;
        memcpy(h + t_off, block_off_host, 4 * size_t(n_blocks + 1));
        if (src_pos_host) memcpy(h + t_pos, src_pos_host, 4 * size_t(n));
        memcpy(h + t_key, key_host, sizeof(uint32_t) * kN);
        CNA_CUDA(cudaMemcpyAsync(tables, h, t_bytes, cudaMemcpyHostToDevice, st));
        CNA_CUDA(cudaMemcpyAsync(raw, h + t_key, sizeof(uint32_t) * kN, cudaMemcpyHostToDevice, st));
        CNA_CUDA(cudaEventRecord(g_stage.copied, st));
    }
    const int64_t *blk_dev0 = reinterpret_cast<const int64_t *>(tables + t_dev0);
    const int32_t *block_off = reinterpret_cast<const int32_t *>(tables + t_off);
    const int32_t *src_pos = src_pos_host ? reinterpret_cast<const int32_t *>(tables + t_pos) : nullptr;
    CNA_CUDA(cudaMemsetAsync(ambiguous, 0, sizeof(int), st));
    CNA_CUDA(cudaMemsetAsync(tail, 0, 4 * sizeof(double), st));
    if (blocks > 0) {
        mt_stream_kernel<<<1, 256, 0, st>>>(raw, blocks);
        CNA_LAUNCHED("mt_stream_kernel");
    }
    const uint32_t *stream_words = raw + pos;
    if (n_pairs > 0) {
        attempts_count_kernel<<<unsigned(n_chunks), 256, 0, st>>>(stream_words, n_att, chunk_count);
        CNA_LAUNCHED("attempts_count_kernel");
    }
    chunk_scan_kernel<<<1, 1024, 0, st>>>(chunk_count, n_pairs > 0 ? n_chunks : 0, chunk_off);
    CNA_LAUNCHED("chunk_scan_kernel");
    Emit e;
    e.stream = stream_words;
    e.n_att = n_att;
    e.chunk_off = chunk_off;
    e.first = first;
    e.count = count;
    e.n_pairs = n_pairs;
    e.dev = dev;
    e.blk_dev0 = blk_dev0;
    e.block_off = block_off;
    e.n_blocks = n_blocks;
    e.num = num;
    e.n = n;
    e.tail = tail;
    if (first)  // legacy_gauss hands out the cached deviate first: block 0, row 0, permutation 0
        CNA_CUDA(cudaMemcpyAsync(dev, tables + t_gauss, sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (n_pairs > 0) {
        attempts_emit_kernel<<<unsigned(n_chunks), 256, 0, st>>>(e);
        CNA_LAUNCHED("attempts_emit_kernel");
    }
    final_state_kernel<<<1, 256, 0, st>>>(raw, pos, tail, chunk_off, n_pairs > 0 ? n_chunks : 0, n_pairs, state_out);
    CNA_LAUNCHED("final_state_kernel");
    const size_t smem = sizeof(double) * 16 * size_t(n);
    CNA_CUDA(cudaFuncSetAttribute(rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int64_t want = (num + 7) / 8, cap = int64_t(num_sms()) * 4;
    rank_kernel<<<unsigned(want < cap ? want : cap), 256, smem, st>>>(dev, num, int(n), n_blocks, block_off, src_pos, out,
                                                                    ld_out, ambiguous);
    CNA_LAUNCHED("rank_kernel");
    return CNA_OK;
}
