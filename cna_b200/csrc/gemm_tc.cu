// tcgen05 (5th-generation tensor core) versions of the dense contractions of the path:
//
//   xb_tc_kernel<HIST>    histogram of (X . Ycond / n)^2 against the threshold edges
//                         (_association.py:99 + _stats.py:52-54)
//   xb_tc_kernel<STORE>   out = X . B   (_nam.py:106, V = NAM^T U / sqrt(svs))
//   gram_tc_kernel        G += X^T X    (_nam.py:105)
//
// Precision.  The reference is fp64 and the north-star tolerance is 1e-5, so a single low-precision
// pass is not enough.  Operands are split once into two fp16 planes, x = hi + lo with
// hi = fp16(x), lo = fp16(x - hi) (22 significant bits; the dropped lo.lo term is 2^-22 relative),
// and every contraction is three kind::f16 MMAs  lo.hi + hi.lo + hi.hi  accumulated in fp32 in
// TMEM.  That costs 1.5x a single TF32 pass (kind::f16 consumes K=16 per instruction, TF32 K=8),
// half of a 3xTF32 split, and the two planes together are exactly as many bytes as the fp32 matrix.
//
// Structure (both kernels): persistent CTAs, one per SM; warp 0 = TMA producer (one elected lane),
// warp 1 = TMEM allocator + MMA issuer (one elected lane), remaining warps = epilogue
// (tcgen05.ld -> registers).  smem ring of operand stages guarded by full/empty mbarriers;
// accumulators in TMEM guarded by tmem_full/tmem_empty mbarriers.
//
//   * xb: A = X rows (M = 128 cells, K = samples contiguous: K-major), B = Ycond^T rows
//     (N = 256 permutations, K-major).  One k-step (16 samples) per stage, SWIZZLE_32B boxes.
//     Two 128x256 fp32 accumulators (all 512 TMEM columns) so the histogram epilogue of one tile
//     overlaps the MMAs of the next.
//   * gram: the contraction runs over cells, so both operands are MN-major views of the same smem
//     tile of X (samples contiguous): boxes of 64 samples x 32 cells, SWIZZLE_128B.  The fp32
//     accumulators (2 m-tiles x n columns) are flushed to fp64 every 512 cells: fp32 tensor-core
//     accumulation truncates, and an unflushed sum over 10^4 steps would bias the diagonal by more
//     than the 1e-5 budget.  Partial Grams live in a per-CTA fp64 scratch (no atomics; the final
//     reduction is deterministic).
#include <atomic>

#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace cna {
namespace tc {

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// A wait that can never hang the GPU: after ~2 s of SM clock the kernel traps (the launch fails
// with an error instead of spinning until the watchdog).  A healthy wait is microseconds.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] . B[smem]^T, fp16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once every MMA issued so far by this thread has completed (implies
// tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address [0,14), leading byte
// offset [16,30), stride byte offset [32,46) (all >> 4), version = 1 at [46,48), layout type at
// [61,64) (2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= uint64_t((addr >> 4) & 0x3FFF);
    d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(layout) << 61;
    return d;
}
// Instruction descriptor for kind::f16 (cute::UMMA::InstrDescriptor): D = fp32 (bit 4), A/B = fp16
// (format 0), a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major), N >> 3 at [17,23),
// M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t instr_desc_f16(int m, int n, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (uint32_t(a_mn_major) << 15) | (uint32_t(b_mn_major) << 16) | (uint32_t(n >> 3) << 17) |
           (uint32_t(m >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// fp16 hi/lo split
// ---------------------------------------------------------------------------------------------
__global__ void split_f16_kernel(const float *__restrict__ src, int64_t ld_src, int64_t src_rows, int src_cols,
                                 int transpose, __half *__restrict__ hi, __half *__restrict__ lo, int64_t ld_dst,
                                 int64_t dst_rows) {
    int64_t total = dst_rows * ld_dst;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        int64_t r = i / ld_dst, c = i - r * ld_dst;
        int64_t sr = transpose ? c : r, sc = transpose ? r : c;
        float v = (sr < src_rows && sc < src_cols) ? src[sr * ld_src + sc] : 0.f;
        __half h = __float2half_rn(v);
        hi[i] = h;
        lo[i] = __float2half_rn(v - __half2float(h));
    }
}

// ---------------------------------------------------------------------------------------------
// X . B^T  (cells x samples) . (samples x outputs), histogram or store epilogue
// ---------------------------------------------------------------------------------------------
constexpr int kBM = 128;                 // cells per tile (UMMA M)
constexpr int kBN = 256;                 // output columns per tile (UMMA N)
constexpr int kStepBytesA = kBM * 32;    // one k-step (16 fp16) of one plane of A
constexpr int kStepBytesB = kBN * 32;
constexpr int kStageBytes = 2 * kStepBytesA + 2 * kStepBytesB;  // hi+lo of A and B: 24 KiB
constexpr int kStages = 7;
constexpr int kXbEpiWarps = 16;          // four warps per TMEM lane quarter, each owning a quarter of the columns
constexpr int kXbThreads = 64 + kXbEpiWarps * 32;
constexpr int kQueue = 512;              // candidates of one 16-column batch of one warp (32 lanes x 16)
constexpr uint32_t kSw32 = 6, kSw128 = 2;

enum class Epi { STORE, HIST };

struct XbTcArgs {
    int64_t n_rows;
    int n_ksteps;
    int n_out;
    // STORE
    float *out;
    int64_t ld_out;
    // HIST
    const double *edges;
    int n_edges;               // number of edges; with n_edges_dev: capacity of the tables
    const int *n_edges_dev;    // optional: the count left on the device by cna_fdr_thresholds
    unsigned long long *hist;  // [n_edges], summed over all output columns
    double inv_n;
    float reject_below;        // with n_edges_dev: derived from edges[0] inside the kernel
    double reject_scale;       // n^2 (1 - 1e-5)
};

template <Epi E>
__global__ void __launch_bounds__(kXbThreads, 1)
xb_tc_kernel(const __grid_constant__ CUtensorMap tm_ah, const __grid_constant__ CUtensorMap tm_al,
             const __grid_constant__ CUtensorMap tm_bh, const __grid_constant__ CUtensorMap tm_bl, XbTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + kStages * kStageBytes);
    uint64_t *empty = full + kStages;
    uint64_t *tfull = empty + kStages;
    uint64_t *tempty = tfull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
    float *guess = reinterpret_cast<float *>(tmem_slot + 2);              // [2]: sqrt(e_0), bins per unit sqrt
    float *queue_v = guess + 2;                                            // [warps][kQueue]
    uint32_t *hist_s = reinterpret_cast<uint32_t *>(queue_v + kXbEpiWarps * kQueue);  // [n_edges] CTA-private
    float *elo_s = reinterpret_cast<float *>(hist_s + a.n_edges);  // [n_edges] edge b, nudged up
    float *ehi_s = elo_s + a.n_edges;                               // [n_edges] edge b+1, nudged down
    double *edges_s = reinterpret_cast<double *>(ehi_s + a.n_edges + (a.n_edges & 1));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_row_tiles = int((a.n_rows + kBM - 1) / kBM);
    const int n_chunks = (a.n_out + kBN - 1) / kBN;
    int n_edges = a.n_edges;
    float reject_below = a.reject_below;
    if (E == Epi::HIST && a.n_edges_dev) {
        n_edges = min(__ldg(a.n_edges_dev), a.n_edges);
        const double rb = (n_edges > 0 ? __ldg(a.edges) : 1e300) * a.reject_scale;
        reject_below = rb > 0.0 ? float(rb) * (1.0f - 1e-6f) : 0.f;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull + b, 1);
            mbar_init(tempty + b, kXbEpiWarps);  // one arrive per epilogue warp
        }
        fence_barrier_init();
        tma_prefetch_desc(&tm_ah);
        tma_prefetch_desc(&tm_al);
        tma_prefetch_desc(&tm_bh);
        tma_prefetch_desc(&tm_bl);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    if (E == Epi::HIST) {
        for (int t = threadIdx.x; t < n_edges; t += blockDim.x) {
            edges_s[t] = a.edges[t];
            hist_s[t] = 0;
            // fp32 copies with a 2e-6 relative guard band (the fp32 value of z^2 carries < 3e-7 of
            // rounding): a product strictly inside (elo[b], ehi[b]) is in bin b beyond doubt
            elo_s[t] = float(a.edges[t] * (1.0 + 2e-6));
            ehi_s[t] = (t + 1 < n_edges) ? float(a.edges[t + 1] * (1.0 - 2e-6)) : 3.0e38f;
        }
        if (threadIdx.x == 32 && n_edges > 0) {
            // the thresholds are (close to) an arithmetic progression, so sqrt(edge) is close to
            // linear in the bin index: a one-multiply first guess, corrected against the table
            double s0 = sqrt(fmax(a.edges[0], 0.0)), s1 = sqrt(fmax(a.edges[n_edges - 1], 0.0));
            guess[0] = float(s0);
            guess[1] = (s1 > s0) ? float((n_edges - 1) / (s1 - s0)) : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer ----
            int stage = 0;
            uint32_t phase = 0;
            for (int rt = blockIdx.x; rt < n_row_tiles; rt += gridDim.x)
                for (int nc = 0; nc < n_chunks; ++nc)
                    for (int k = 0; k < a.n_ksteps; ++k) {
                        mbar_wait(empty + stage, phase ^ 1);
                        uint8_t *st = smem + stage * kStageBytes;
                        mbar_expect_tx(full + stage, kStageBytes);
                        tma_load_2d(st, &tm_ah, k * 16, rt * kBM, full + stage);
                        tma_load_2d(st + kStepBytesA, &tm_al, k * 16, rt * kBM, full + stage);
                        tma_load_2d(st + 2 * kStepBytesA, &tm_bh, k * 16, nc * kBN, full + stage);
                        tma_load_2d(st + 2 * kStepBytesA + kStepBytesB, &tm_bl, k * 16, nc * kBN, full + stage);
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer ----
            constexpr uint32_t idesc = instr_desc_f16(kBM, kBN, 0, 0);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int rt = blockIdx.x; rt < n_row_tiles; rt += gridDim.x)
                for (int nc = 0; nc < n_chunks; ++nc) {
                    mbar_wait(tempty + acc, acc_phase ^ 1);  // epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t d = tmem_base + uint32_t(acc * kBN);
                    for (int k = 0; k < a.n_ksteps; ++k) {
                        mbar_wait(full + stage, phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + stage * kStageBytes);
                        // K-major, SWIZZLE_32B: 8-row groups are 256 B apart, one k-step per box
                        const uint64_t ah = smem_desc(sa, 16, 256, kSw32);
                        const uint64_t al = smem_desc(sa + kStepBytesA, 16, 256, kSw32);
                        const uint64_t bh = smem_desc(sa + 2 * kStepBytesA, 16, 256, kSw32);
                        const uint64_t bl = smem_desc(sa + 2 * kStepBytesA + kStepBytesB, 16, 256, kSw32);
                        umma_f16(d, al, bh, idesc, k > 0);  // small terms first
                        umma_f16(d, ah, bl, idesc, 1);
                        umma_f16(d, ah, bh, idesc, 1);
                        umma_commit(empty + stage);  // smem slot free once these MMAs have read it
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    umma_commit(tfull + acc);  // accumulator complete
                    if (++acc == 2) {
                        acc = 0;
                        acc_phase ^= 1;
                    }
                }
        }
    } else {  // ---- epilogue: warps 2..17; TMEM lane quarter = warp % 4, column quarter = (warp - 2) / 4 ----
        constexpr int kColsPerWarp = kBN / (kXbEpiWarps / 4);
        const int q = warp & 3, part = (warp - 2) >> 2, ew = warp - 2;
        float *qv = queue_v + ew * kQueue;
        const float g0 = (E == Epi::HIST) ? guess[0] : 0.f, g1 = (E == Epi::HIST) ? guess[1] : 0.f;
        const float inv_n_f = float(a.inv_n);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int rt = blockIdx.x; rt < n_row_tiles; rt += gridDim.x) {
            const int64_t row = int64_t(rt) * kBM + q * 32 + lane;
            const bool row_ok = row < a.n_rows;
            for (int nc = 0; nc < n_chunks; ++nc) {
                mbar_wait(tfull + acc, acc_phase);
                tc_fence_after();
                const uint32_t t0 = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * kBN);
                for (int c0 = part * kColsPerWarp; c0 < (part + 1) * kColsPerWarp; c0 += 16) {
                    const int col0 = nc * kBN + c0;
                    if (col0 >= a.n_out) break;  // uniform over the warp
                    uint32_t r[16];
                    tmem_ld16(t0 + c0, r);
                    tmem_ld_wait();
                    if (E == Epi::STORE) {
                        if (row_ok) {
                            float *o = a.out + row * a.ld_out + col0;
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (col0 + j < a.n_out) o[j] = __uint_as_float(r[j]);
                        }
                    } else {
                        // pass 1: which of my 16 products can reach the first edge at all?
                        uint32_t mask = 0;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float v = __uint_as_float(r[j]);
                            if (v * v >= reject_below && col0 + j < a.n_out) mask |= 1u << j;
                        }
                        if (!row_ok) mask = 0;
                        // compact the survivors of the whole warp into its queue
                        int cnt = __popc(mask), off = cnt;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            int t = __shfl_up_sync(kFull, off, o);
                            if (lane >= o) off += t;
                        }
                        const int total = __shfl_sync(kFull, off, 31);
                        if (total == 0) continue;
                        off -= cnt;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (mask & (1u << j)) qv[off++] = __uint_as_float(r[j]);
                        __syncwarp();
                        // pass 2: all lanes busy on survivors: bin lookup + one shared-memory atomic each
                        // (the FDR only needs the histogram summed over the null columns, _stats.py:79-80)
                        for (int i = lane; i < total; i += 32) {
                            const float vf = qv[i];
                            const float zf = vf * inv_n_f, z2f = zf * zf;
                            int b = int((sqrtf(z2f) - g0) * g1);
                            b = max(0, min(b, n_edges - 1));
                            if (!(z2f > elo_s[b] && z2f < ehi_s[b])) {
                                // within rounding distance of an edge (or a missed guess): decide in fp64
                                const double z = double(vf) * a.inv_n;
                                const double z2 = z * z;
                                if (!(z2 >= edges_s[0])) continue;
                                while (b + 1 < n_edges && edges_s[b + 1] <= z2) ++b;  // largest b with edges[b] <= z2
                                while (b > 0 && edges_s[b] > z2) --b;
                            }
                            atomicAdd(hist_s + b, 1u);
                        }
                        __syncwarp();
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty + acc);
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
    if (E == Epi::HIST)  // a CTA sees < 2^32 products per bin between flushes (rows/148 x columns)
        for (int t = threadIdx.x; t < n_edges; t += blockDim.x)
            if (hist_s[t]) atomicAdd(a.hist + t, static_cast<unsigned long long>(hist_s[t]));
}

// ---------------------------------------------------------------------------------------------
// Gram: G += X^T X, contraction over cells
// ---------------------------------------------------------------------------------------------
// Output tiles: m-tile t = samples [128t, 128t+128) x columns [128t, n16) (the blocks below the
// diagonal are transposes of blocks above it).  TMEM has 512 fp32 columns per SM, so the CTAs are
// split into groups; a group owns consecutive m-tiles whose widths sum to <= 512 columns
// (n <= 256: one group {0,1}, every CTA reads X once; n = 500: groups {0}, {1}, {2,3}) and a share
// of the CTAs proportional to its MMA work, and streams only the 64-sample boxes its tiles touch.
//
// fp32 accumulation in TMEM truncates (measured: -3.2e-8 relative per accumulation step on an
// all-positive sum), so the accumulators are flushed into fp64 every kGChunkStages * 32 = 256 cells
// (48 steps, bias ~1.5e-6 on the diagonal, inside the 1e-5 budget with margin).
constexpr int kGK = 32;                       // cells per stage (two k-steps)
constexpr int kGBoxBytes = kGK * 128;         // 64 samples x 32 cells of fp16
constexpr int kGRingBytes = 192 * 1024;       // operand ring: 6 stages of 4 boxes ... 3 stages of 8
constexpr int kGMaxStages = 6;
constexpr int kGChunkStages = 8;
constexpr int kGEpiWarps = 16;
constexpr int kGThreads = 64 + kGEpiWarps * 32;
constexpr int kGMaxGroups = 4;

struct GramGroup {
    int mt0, n_mt;      // m-tiles [mt0, mt0 + n_mt)
    int box0, n_boxes;  // 64-sample boxes [box0, box0 + n_boxes) streamed by this group
    int width;          // TMEM / scratch columns: sum over its tiles of (n16 - 128 t)
    int cta0, n_ctas;   // CTAs [cta0, cta0 + n_ctas)
    int64_t scratch_off;  // doubles
};

struct GramTcArgs {
    int64_t n_rows;
    int n16;        // samples rounded up to 16
    int n_groups;
    GramGroup g[kGMaxGroups];
    double *scratch;  // per CTA: [width of its group][128]
};

__global__ void __launch_bounds__(kGThreads, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_l, GramTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    int gi = 0;
    while (gi + 1 < a.n_groups && int(blockIdx.x) >= a.g[gi + 1].cta0) ++gi;
    const GramGroup grp = a.g[gi];
    const int local = int(blockIdx.x) - grp.cta0;  // this CTA takes chunks local, local + n_ctas, ...
    const int plane_bytes = grp.n_boxes * kGBoxBytes;
    const int stage_bytes = 2 * plane_bytes;
    const int n_stages_ring = min(kGMaxStages, kGRingBytes / stage_bytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + kGRingBytes);
    uint64_t *empty = full + kGMaxStages;
    uint64_t *tfull = empty + kGMaxStages;
    uint64_t *tempty = tfull + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_stages_total = (a.n_rows + kGK - 1) / kGK;
    const int64_t n_chunks = (n_stages_total + kGChunkStages - 1) / kGChunkStages;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGMaxStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(tfull, 1);
        mbar_init(tempty, kGEpiWarps);
        fence_barrier_init();
        tma_prefetch_desc(&tm_h);
        tma_prefetch_desc(&tm_l);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer ----
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t ch = local; ch < n_chunks; ch += grp.n_ctas) {
                int64_t s0 = ch * kGChunkStages, s1 = s0 + kGChunkStages;
                if (s1 > n_stages_total) s1 = n_stages_total;
                for (int64_t s = s0; s < s1; ++s) {
                    mbar_wait(empty + stage, phase ^ 1);
                    uint8_t *st = smem + stage * stage_bytes;
                    mbar_expect_tx(full + stage, uint32_t(stage_bytes));
                    for (int b = 0; b < grp.n_boxes; ++b) {
                        const int c0 = (grp.box0 + b) * 64;
                        tma_load_2d(st + b * kGBoxBytes, &tm_h, c0, int(s * kGK), full + stage);
                        tma_load_2d(st + plane_bytes + b * kGBoxBytes, &tm_l, c0, int(s * kGK), full + stage);
                    }
                    if (++stage == n_stages_ring) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer ----
            int stage = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int64_t ch = local; ch < n_chunks; ch += grp.n_ctas) {
                int64_t s0 = ch * kGChunkStages, s1 = s0 + kGChunkStages;
                if (s1 > n_stages_total) s1 = n_stages_total;
                mbar_wait(tempty, acc_phase ^ 1);
                tc_fence_after();
                for (int64_t s = s0; s < s1; ++s) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
                    const uint32_t sh = smem_u32(smem + stage * stage_bytes), sl = sh + plane_bytes;
#pragma unroll
                    for (int ks = 0; ks < kGK / 16; ++ks) {
                        // MN-major, SWIZZLE_128B: 64-sample atoms are one box (LBO) apart, 8-cell
                        // groups 1024 B (SBO) apart; a k-step of 16 cells is 2048 B
                        const uint32_t koff = ks * 2048;
                        const uint32_t first = (s == s0 && ks == 0) ? 0u : 1u;
                        uint32_t col = 0;  // TMEM column of the current tile
                        for (int t = 0; t < grp.n_mt; ++t) {
                            const int mt = grp.mt0 + t, width = a.n16 - 128 * mt;
                            const uint32_t aoff = uint32_t(2 * mt - grp.box0) * kGBoxBytes;
                            const uint64_t ah = smem_desc(sh + aoff + koff, kGBoxBytes, 1024, kSw128);
                            const uint64_t al = smem_desc(sl + aoff + koff, kGBoxBytes, 1024, kSw128);
                            // B = columns [128 mt, n16): the same boxes as A onwards, at most 256
                            // columns (4 boxes) per instruction
                            for (int c0 = 0; c0 < width; c0 += 256) {
                                const int nn = min(256, width - c0);
                                const uint32_t boff = aoff + uint32_t(c0 / 64) * kGBoxBytes;
                                const uint64_t bh = smem_desc(sh + boff + koff, kGBoxBytes, 1024, kSw128);
                                const uint64_t bl = smem_desc(sl + boff + koff, kGBoxBytes, 1024, kSw128);
                                const uint32_t idesc = instr_desc_f16(128, nn, 1, 1);
                                const uint32_t d = tmem_base + col + uint32_t(c0);
                                umma_f16(d, al, bh, idesc, first);  // small terms first
                                umma_f16(d, ah, bl, idesc, 1);
                                umma_f16(d, ah, bh, idesc, 1);
                            }
                            col += uint32_t(width);
                        }
                    }
                    umma_commit(empty + stage);
                    if (++stage == n_stages_ring) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(tfull);
                acc_phase ^= 1;
            }
        }
    } else {  // ---- epilogue: 16 warps; lane quarter = warp % 4, column groups dealt round-robin ----
        const int q = warp & 3, part = (warp - 2) >> 2;
        const int n_colgroups = grp.width / 16;
        double *mine = a.scratch + grp.scratch_off + int64_t(local) * grp.width * 128;
        uint32_t acc_phase = 0;
        bool first = true;
        for (int64_t ch = local; ch < n_chunks; ch += grp.n_ctas) {
            mbar_wait(tfull, acc_phase);
            tc_fence_after();
            for (int g = part; g < n_colgroups; g += kGEpiWarps / 4) {
                uint32_t r[16];
                tmem_ld16(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(g * 16), r);
                double *p = mine + int64_t(g * 16) * 128 + q * 32 + lane;
                if (first) {
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) p[j * 128] = double(__uint_as_float(r[j]));
                } else {
                    double old[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) old[j] = p[j * 128];
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) p[j * 128] = old[j] + double(__uint_as_float(r[j]));
                }
            }
            first = false;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty);
            acc_phase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// G[i][j] += sum over the CTAs of the owning group of their partial (fixed order: deterministic).
// Entries below the diagonal blocks are read from their mirror image.
__global__ void gram_reduce_kernel(GramTcArgs a, int64_t n_chunks, int n, double *__restrict__ gram) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    int i = idx / n, j = idx - i * n;
    if (j < (i & ~127)) {  // below the diagonal block of row i: mirror
        int t = i;
        i = j;
        j = t;
    }
    const int mt = i >> 7;
    int gi = 0;
    while (gi + 1 < a.n_groups && mt >= a.g[gi + 1].mt0) ++gi;
    const GramGroup grp = a.g[gi];
    int col = j - 128 * mt;  // column inside tile mt, then inside the group
    for (int t = grp.mt0; t < mt; ++t) col += a.n16 - 128 * t;
    const int64_t per_cta = int64_t(grp.width) * 128;
    const double *base = a.scratch + grp.scratch_off + int64_t(col) * 128 + (i & 127);
    const int live = int(n_chunks < grp.n_ctas ? n_chunks : grp.n_ctas);  // CTAs that got a chunk
    double acc = 0.0;
    for (int b = 0; b < live; ++b) acc += base[b * per_cta];
    gram[idx] += acc;
}

// Groups of consecutive m-tiles (greedy, <= 512 TMEM columns each) and their CTA shares.
static int gram_plan(int n, int n_ctas, GramTcArgs &a) {
    a.n16 = (n + 15) / 16 * 16;
    const int n_mt = (a.n16 + 127) / 128, n_boxes = (a.n16 + 63) / 64;
    a.n_groups = 0;
    int mt = 0;
    double total_work = 0.0, work[kGMaxGroups];
    while (mt < n_mt) {
        if (a.n_groups == kGMaxGroups) return -1;
        GramGroup &g = a.g[a.n_groups];
        g.mt0 = mt;
        g.n_mt = 0;
        g.width = 0;
        while (mt < n_mt && g.width + (a.n16 - 128 * mt) <= 512) {
            g.width += a.n16 - 128 * mt;
            ++g.n_mt;
            ++mt;
        }
        if (g.n_mt == 0) return -1;
        g.box0 = 2 * g.mt0;
        g.n_boxes = n_boxes - g.box0;
        work[a.n_groups] = double(g.width);
        total_work += work[a.n_groups];
        ++a.n_groups;
    }
    if (n_ctas < a.n_groups) return -1;
    int assigned = 0;
    int64_t off = 0;
    for (int k = 0; k < a.n_groups; ++k) {
        GramGroup &g = a.g[k];
        int share = (k + 1 == a.n_groups) ? n_ctas - assigned : int(n_ctas * work[k] / total_work + 0.5);
        if (share < 1) share = 1;
        if (share > n_ctas - assigned - (a.n_groups - 1 - k)) share = n_ctas - assigned - (a.n_groups - 1 - k);
        g.cta0 = assigned;
        g.n_ctas = share;
        g.scratch_off = off;
        assigned += share;
        off += int64_t(share) * g.width * 128;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp16 tensor [rows][ld] viewed as (inner = cols, outer = rows); out-of-bounds reads give zeros.
static int make_map(CUtensorMap *map, const void *base, int64_t cols, int64_t rows, int64_t ld, int box_cols,
                    int box_rows, CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(CNA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
    cuuint64_t strides[1] = {cuuint64_t(ld) * 2};
    cuuint32_t box[2] = {cuuint32_t(box_cols), cuuint32_t(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return set_error(CNA_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", int(rc));
    return CNA_OK;
}

static std::atomic<int> g_cta_cap{0};

static int xb_tc_launch(Epi epi, const void *xh, const void *xl, int64_t ld16, int64_t n_rows, int n, const void *bh,
                        const void *bl, int64_t ld16_b, int n_out, XbTcArgs a, cudaStream_t st) {
    CNA_REQUIRE(n_rows > 0 && n_rows < (int64_t(1) << 31) && n > 0 && n_out > 0, "xb_tc: bad shape");
    CNA_REQUIRE(ld16 % 8 == 0 && ld16_b % 8 == 0 && ld16 >= n && ld16_b >= n, "xb_tc: plane leading dimensions must be multiples of 8 and >= n");
    CNA_REQUIRE(((reinterpret_cast<uintptr_t>(xh) | reinterpret_cast<uintptr_t>(xl) | reinterpret_cast<uintptr_t>(bh) |
                  reinterpret_cast<uintptr_t>(bl)) & 15) == 0, "xb_tc: planes must be 16-byte aligned");
    CUtensorMap tm_ah, tm_al, tm_bh, tm_bl;
    int rc;
    if ((rc = make_map(&tm_ah, xh, n, n_rows, ld16, 16, kBM, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
    if ((rc = make_map(&tm_al, xl, n, n_rows, ld16, 16, kBM, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
    if ((rc = make_map(&tm_bh, bh, n, n_out, ld16_b, 16, kBN, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
    if ((rc = make_map(&tm_bl, bl, n, n_out, ld16_b, 16, kBN, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
    a.n_rows = n_rows;
    a.n_ksteps = (n + 15) / 16;
    a.n_out = n_out;
    size_t smem = size_t(kStages) * kStageBytes + 1024 + 256 + size_t(kXbEpiWarps) * kQueue * 4 +
                  (epi == Epi::HIST ? (sizeof(double) + sizeof(uint32_t) + 2 * sizeof(float)) * (a.n_edges + 2) : 0);
    int64_t tiles = (n_rows + kBM - 1) / kBM;
    int ctas = num_sms();
    const int cap = g_cta_cap.load(std::memory_order_relaxed);
    if (cap > 0 && cap < ctas) ctas = cap;  // the caller keeps SMs free for a kernel on another stream
    unsigned grid = unsigned(tiles < ctas ? tiles : ctas);
    if (epi == Epi::HIST) {
        CNA_CUDA(cudaFuncSetAttribute(xb_tc_kernel<Epi::HIST>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        xb_tc_kernel<Epi::HIST><<<grid, kXbThreads, smem, st>>>(tm_ah, tm_al, tm_bh, tm_bl, a);
        CNA_LAUNCHED("xb_tc_kernel<HIST>");
    } else {
        CNA_CUDA(cudaFuncSetAttribute(xb_tc_kernel<Epi::STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        xb_tc_kernel<Epi::STORE><<<grid, kXbThreads, smem, st>>>(tm_ah, tm_al, tm_bh, tm_bl, a);
        CNA_LAUNCHED("xb_tc_kernel<STORE>");
    }
    return CNA_OK;
}

}  // namespace tc
}  // namespace cna

using namespace cna;
using namespace cna::tc;

extern "C" {

int cna_split_f16(const float *src, int64_t ld_src, int64_t src_rows, int src_cols, int transpose, void *hi,
                  void *lo, int64_t ld_dst, int64_t dst_rows, void *stream) {
    CNA_REQUIRE(src_rows >= 0 && src_cols >= 0 && ld_dst > 0 && dst_rows >= 0 && hi && lo, "cna_split_f16: bad arguments");
    CNA_REQUIRE(transpose ? (dst_rows >= src_cols && ld_dst >= src_rows) : (dst_rows >= src_rows && ld_dst >= src_cols),
                "cna_split_f16: destination smaller than source");
    int64_t total = dst_rows * ld_dst;
    if (total == 0) return CNA_OK;
    int64_t blocks = (total + 255) / 256;
    int64_t cap = int64_t(num_sms()) * 16;
    split_f16_kernel<<<unsigned(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(
        src, ld_src, src_rows, src_cols, transpose, static_cast<__half *>(hi), static_cast<__half *>(lo), ld_dst, dst_rows);
    CNA_LAUNCHED("split_f16_kernel");
    return CNA_OK;
}

int cna_right_multiply_tc(const void *xh, const void *xl, int64_t ld16, int64_t n_rows, int n, const void *bth,
                          const void *btl, int64_t ld16_b, int n_out, float *out, int64_t ld_out, void *stream) {
    CNA_REQUIRE(out && ld_out >= n_out, "cna_right_multiply_tc: bad output");
    if (n_rows == 0) return CNA_OK;
    XbTcArgs a{};
    a.out = out;
    a.ld_out = ld_out;
    return xb_tc_launch(Epi::STORE, xh, xl, ld16, n_rows, n, bth, btl, ld16_b, n_out, a, as_stream(stream));
}

int cna_null_hist_tc_dev(const void *xh, const void *xl, int64_t ld16, int64_t n_rows, int n, const void *yth,
                         const void *ytl, int64_t ld16_y, int n_null, const double *edges, int n_edges,
                         const int32_t *n_edges_dev, double edge0, uint64_t *hist, void *stream) {
    CNA_REQUIRE(n_edges > 0 && n_edges <= 1024, "cna_null_hist_tc: 1..1024 edges supported (got %d)", n_edges);
    CNA_REQUIRE(n_rows * int64_t(n_null) / 64 < (int64_t(1) << 32), "cna_null_hist_tc: too many products per CTA");
    CNA_REQUIRE(edges && hist, "cna_null_hist_tc: null pointer");
    if (n_rows == 0 || n_null == 0) return CNA_OK;
    XbTcArgs a{};
    a.edges = edges;
    a.n_edges = n_edges;
    a.n_edges_dev = n_edges_dev;
    a.hist = reinterpret_cast<unsigned long long *>(hist);
    a.inv_n = 1.0 / double(n);
    a.reject_scale = double(n) * double(n) * (1.0 - 1e-5);
    double rb = edge0 * a.reject_scale;
    a.reject_below = rb > 0.0 ? float(rb) * (1.0f - 1e-6f) : 0.f;
    return xb_tc_launch(Epi::HIST, xh, xl, ld16, n_rows, n, yth, ytl, ld16_y, n_null, a, as_stream(stream));
}

int cna_null_hist_tc(const void *xh, const void *xl, int64_t ld16, int64_t n_rows, int n, const void *yth,
                     const void *ytl, int64_t ld16_y, int n_null, const double *edges, int n_edges, double edge0,
                     uint64_t *hist, void *stream) {
    return cna_null_hist_tc_dev(xh, xl, ld16, n_rows, n, yth, ytl, ld16_y, n_null, edges, n_edges, nullptr, edge0,
                                hist, stream);
}

int64_t cna_gram_tc_workspace(int n) {
    if (n <= 0 || n > 512) return -1;
    GramTcArgs a{};
    if (gram_plan(n, num_sms(), a) != 0) return -1;
    const GramGroup &last = a.g[a.n_groups - 1];
    return (last.scratch_off + int64_t(last.n_ctas) * last.width * 128) * int64_t(sizeof(double));
}

int cna_gram_tc(const void *xh, const void *xl, int64_t ld16, int64_t n_rows, int n, double *gram, void *workspace,
                int64_t workspace_bytes, void *stream) {
    CNA_REQUIRE(n > 0 && n <= 512, "cna_gram_tc: 1..512 samples supported (got %d)", n);
    CNA_REQUIRE(n_rows >= 0 && n_rows < (int64_t(1) << 31) && ld16 % 8 == 0 && ld16 >= n, "cna_gram_tc: bad shape");
    CNA_REQUIRE(workspace && workspace_bytes >= cna_gram_tc_workspace(n), "cna_gram_tc: workspace too small");
    CNA_REQUIRE(((reinterpret_cast<uintptr_t>(xh) | reinterpret_cast<uintptr_t>(xl)) & 15) == 0,
                "cna_gram_tc: planes must be 16-byte aligned");
    if (n_rows == 0) return CNA_OK;
    cudaStream_t st = as_stream(stream);
    GramTcArgs a{};
    CNA_REQUIRE(gram_plan(n, num_sms(), a) == 0, "cna_gram_tc: no tile plan for n=%d", n);
    a.n_rows = n_rows;
    a.scratch = static_cast<double *>(workspace);
    CUtensorMap tm_h, tm_l;
    int rc;
    if ((rc = make_map(&tm_h, xh, n, n_rows, ld16, 64, kGK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_map(&tm_l, xl, n, n_rows, ld16, 64, kGK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    int64_t stages = (n_rows + kGK - 1) / kGK;
    int64_t chunks = (stages + kGChunkStages - 1) / kGChunkStages;
    unsigned grid = unsigned(num_sms());  // every group keeps its CTA range; CTAs without a chunk idle
    size_t smem = size_t(kGRingBytes) + 1024 + 256;
    CNA_CUDA(cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    gram_tc_kernel<<<grid, kGThreads, smem, st>>>(tm_h, tm_l, a);
    CNA_LAUNCHED("gram_tc_kernel");
    int nn = n * n;
    gram_reduce_kernel<<<(nn + 255) / 256, 256, 0, st>>>(a, chunks, n, gram);
    CNA_LAUNCHED("gram_reduce_kernel");
    return CNA_OK;
}

}  // extern "C"

extern "C" int cna_tc_max_ctas(int cap) {
    return tc::g_cta_cap.exchange(cap < 0 ? 0 : cap, std::memory_order_relaxed);
}
