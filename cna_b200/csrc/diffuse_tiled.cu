// Kernel (i), shared-memory-staged form: one diffusion step  out = A'.in + diag*in  with the gathered
// rows staged in shared memory by TMA instead of being pulled through L1 edge by edge.
//
// Reference: src/cna/tools/_nam.py:33 (scipy csr_matvecs).  Same arithmetic, same order of additions per
// row as cna_diffuse_step_f32 (the results are bit-identical): only the route of the operands changes.
//
// Why: with ~38 neighbours per cell every row of the state is gathered ~38 times per step.  Pulled
// through L1 (LDG.128, ~70 B/clk/SM) the step is bound by the L1 data pipe and by 24 GB of L2 -> L1
// traffic (profiles/r01c_ncu_spmm_summary.txt).  Here the rows of the stored cell order are cut into
// TILES of <= 64 consecutive output rows; the plan made when the graph becomes resident holds, per tile,
// the sorted list of DISTINCT source rows its edges reference (<= kCap) and, per edge, the position of its
// source in that list.  A persistent CTA per SM walks (tile, 32-column slab) items:
//   * one producer warp issues cp.async.bulk.tensor ... tile::gather4 copies (4 source rows x 128 B per
//     instruction, completion on an mbarrier) into one of two shared-memory buffers;
//   * 16 consumer warps, a quarter-warp per output row, run the row's edges in CSR order out of shared
//     memory (LDS.128, 128 B/clk/SM) and write the finished 128-byte slab of the output row.
// A source row crosses L2 -> SM once per tile that references it instead of once per edge.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace cna {
namespace tiled {

constexpr int kSlabVec = 8;                   // float4 per slab row (32 floats, 128 B)
constexpr int kSlabBytes = kSlabVec * 16;
constexpr int kTileRows = 64;                 // output rows per tile (one per quarter-warp)
constexpr int kConsumerWarps = kTileRows / 4;
constexpr int kProducerWarps = 8;              // TMA issue is serialised per lane (~15 instructions per copy)
constexpr int kThreads = (kConsumerWarps + kProducerWarps) * 32;
constexpr int kCap = 864;                     // distinct source rows per tile (multiple of 4): 108 KB per buffer
constexpr int kBufBytes = kCap * kSlabBytes;

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
// bounded: a protocol bug fails the launch (trap) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
// four rows (r0..r3) x one box width of columns starting at c0 -> 4 consecutive box rows at dst
__device__ __forceinline__ void tma_gather4(void *dst, const CUtensorMap *map, int c0, int r0, int r1, int r2, int r3,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3),
        "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar) {  // arrives once this thread's copies have landed
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct Args {
    const int32_t *indptr;    // [n_rows + 1]
    const uint2 *epair;       // [nnz]: .x = byte offset of the source row's slab in the tile buffer, .y = weight bits
    const float *diag;        // [n_rows]
    const float *in;
    float *out;
    int64_t ld4;              // leading dimension in float4
    int nvec;                 // float4 per row actually computed
    int n_slabs;
    int64_t in_row_offset;
    const int32_t *tile_row;  // [n_tiles + 1] first output row of each tile
    const int32_t *tile_u;    // [n_tiles + 1] offsets into usrc (multiples of 4)
    const int32_t *usrc;      // distinct source rows of every tile, each list padded to a multiple of 4
    int n_tiles;
    int stage_mode;           // 0: TMA gather4, 1: cp.async (LDGSTS) issued by the producer warp
};

__global__ void __launch_bounds__(kThreads, 1)
spmm_tiled_kernel(const __grid_constant__ CUtensorMap tm_in, Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + 2 * kBufBytes);
    uint64_t *empty = full + 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool cpasync = a.stage_mode == 1;
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(full + b, cpasync ? 32 * kProducerWarps : 1);   // cp.async: every producer lane arrives; TMA: one expect_tx
            mbar_init(empty + b, kConsumerWarps);
        }
        fence_barrier_init();
    }
    __syncthreads();

    if (warp >= kConsumerWarps) {
        // ---------------- producers: warp p takes the groups p, p + P, ... of every item ----------------
        const int pw = warp - kConsumerWarps;
        // The tile's source list is read once into registers (lane l owns the groups l, l + 32, ...: at most
        // kCap / 128 int4 each) and reused for every slab, so no load sits between two copies.
        constexpr int kGroupsPerLane = (kCap / 4 + 32 * kProducerWarps - 1) / (32 * kProducerWarps);
        int it = 0;
        for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
            const int u0 = a.tile_u[t], ug = (a.tile_u[t + 1] - u0) >> 2;  // groups of 4 source rows
            const int4 *src4 = reinterpret_cast<const int4 *>(a.usrc + u0);
            int4 rows[kGroupsPerLane];
#pragma unroll
            for (int i = 0; i < kGroupsPerLane; ++i)
                rows[i] = (lane + 32 * i) * kProducerWarps + pw < ug ? __ldg(src4 + (lane + 32 * i) * kProducerWarps + pw)
                                                                     : make_int4(0, 0, 0, 0);
            for (int s = 0; s < a.n_slabs; ++s, ++it) {
                const int b = it & 1;
                mbar_wait(empty + b, ((it >> 1) & 1) ^ 1);
                uint8_t *buf = smem + b * kBufBytes;
                if (!cpasync) {
                    if (lane == 0 && pw == 0) mbar_expect_tx(full + b, uint32_t(ug) * 4 * kSlabBytes);
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < kGroupsPerLane; ++i) {
                        const int g = (lane + 32 * i) * kProducerWarps + pw;
                        if (g < ug)
                            tma_gather4(buf + g * 4 * kSlabBytes, &tm_in, s * kSlabVec * 4, rows[i].x, rows[i].y,
                                        rows[i].z, rows[i].w, full + b);
                    }
                } else {
                    // 8 lanes per source row, 16 bytes each; columns past the row end are never read back
                    const int sub = lane & 7;
                    const bool in_row = s * kSlabVec + sub < a.ld4;
                    const float4 *base = reinterpret_cast<const float4 *>(a.in) + s * kSlabVec + sub;
#pragma unroll
                    for (int i = 0; i < kGroupsPerLane; ++i) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {  // the group lane j holds, handed round the warp
                            const int g = (j + 32 * i) * kProducerWarps + pw;
                            if (g >= ug) break;  // uniform
                            const int4 r = make_int4(__shfl_sync(kFull, rows[i].x, j), __shfl_sync(kFull, rows[i].y, j),
                                                     __shfl_sync(kFull, rows[i].z, j), __shfl_sync(kFull, rows[i].w, j));
                            const int q = lane >> 3;
                            const int64_t rr = q == 0 ? r.x : (q == 1 ? r.y : (q == 2 ? r.z : r.w));
                            if (in_row) cp_async16(buf + (g * 4 + q) * kSlabBytes + sub * 16, base + rr * a.ld4);
                        }
                    }
                    cp_async_arrive(full + b);
                }
            }
        }
    } else {
        // ---------------- consumers: quarter-warp per output row ----------------
        const int sub = lane & 7, qbase = lane & 24;
        const float4 *in4 = reinterpret_cast<const float4 *>(a.in);
        float4 *out4 = reinterpret_cast<float4 *>(a.out);
        int it = 0;
        for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
            const int row0 = a.tile_row[t], nrows = a.tile_row[t + 1] - row0;
            const int rl = warp * 4 + (lane >> 3);
            const bool has_row = rl < nrows;
            const int64_t row = row0 + (has_row ? rl : 0);
            const int e0 = has_row ? a.indptr[row] : 0, e1 = has_row ? a.indptr[row + 1] : 0;
            int len = e1 - e0;
            int maxlen = len;
            maxlen = max(maxlen, __shfl_xor_sync(kFull, maxlen, 8));
            maxlen = max(maxlen, __shfl_xor_sync(kFull, maxlen, 16));
            const float d = has_row ? a.diag[row] : 0.f;
            // The (position, weight) pairs of the row's first 64 edges live in registers for all slabs of
            // the tile: lane `sub` of the quarter holds edges sub, sub + 8, ...; they are handed round the
            // quarter with shuffles.  Longer rows read the rest from memory in every slab.
            uint2 pr[8];
#pragma unroll
            for (int c = 0; c < 8; ++c)
                pr[c] = c * 8 + sub < len ? __ldg(a.epair + e0 + c * 8 + sub) : make_uint2(0u, 0u);
            for (int s = 0; s < a.n_slabs; ++s, ++it) {
                const int b = it & 1;
                mbar_wait(full + b, (it >> 1) & 1);
                const uint8_t *buf = smem + b * kBufBytes + sub * 16;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                // the edges of the row in CSR order (past the end of a shorter row of the warp: weight 0 on
                // the tile's first source row, skipped)
#define CNA_TILED_EDGE(PAIR, K)                                                           \
    {                                                                                     \
        const uint32_t off = __shfl_sync(kFull, (PAIR).x, qbase + j);                     \
        const float w = __uint_as_float(__shfl_sync(kFull, (PAIR).y, qbase + j));         \
        const float4 x = *reinterpret_cast<const float4 *>(buf + off);                    \
        if ((K) < len) {                                                                  \
            acc.x = fmaf(w, x.x, acc.x);                                                  \
            acc.y = fmaf(w, x.y, acc.y);                                                  \
            acc.z = fmaf(w, x.z, acc.z);                                                  \
            acc.w = fmaf(w, x.w, acc.w);                                                  \
        }                                                                                 \
    }
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c * 8 < maxlen) {  // uniform over the warp
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (c * 8 + j < maxlen) CNA_TILED_EDGE(pr[c], c * 8 + j)
                    }
                }
                for (int k0 = 64; k0 < maxlen; k0 += 8) {
                    const uint2 mine = k0 + sub < len ? __ldg(a.epair + e0 + k0 + sub) : make_uint2(0u, 0u);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (k0 + j < maxlen) CNA_TILED_EDGE(mine, k0 + j)
                }
#undef CNA_TILED_EDGE
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + b);  // this warp is done reading the buffer
                const int c = s * kSlabVec + sub;
                if (has_row && c < a.nvec) {
                    const float4 x = __ldg(in4 + (row + a.in_row_offset) * a.ld4 + c);  // self term last (_nam.py:33)
                    acc.x = fmaf(d, x.x, acc.x);
                    acc.y = fmaf(d, x.y, acc.y);
                    acc.z = fmaf(d, x.z, acc.z);
                    acc.w = fmaf(d, x.w, acc.w);
                    out4[row * a.ld4 + c] = acc;
                }
            }
        }
    }
}

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

}  // namespace tiled
}  // namespace cna

using namespace cna;
using namespace cna::tiled;

extern "C" {

int cna_diffuse_tile_limits(int32_t *tile_rows, int32_t *tile_sources) {
    if (tile_rows) *tile_rows = kTileRows;
    if (tile_sources) *tile_sources = kCap;
    return CNA_OK;
}

int cna_diffuse_step_f32_tiled(const int32_t *indptr, const void *epair, const float *diag, const float *in,
                               float *out, int64_t n_rows, int64_t in_rows, int n_cols, int64_t ld,
                               int64_t in_row_offset, const int32_t *tile_row, const int32_t *tile_u,
                               const int32_t *usrc, int n_tiles, int stage_mode, void *stream) {
    CNA_REQUIRE(n_rows >= 0 && n_cols > 0 && ld >= n_cols && ld % 4 == 0 && in_rows > 0,
                "cna_diffuse_step_f32_tiled: bad shape");
    CNA_REQUIRE(in != out, "cna_diffuse_step_f32_tiled: in-place step is not supported");
    CNA_REQUIRE(indptr && epair && diag && tile_row && tile_u && usrc && n_tiles >= 0,
                "cna_diffuse_step_f32_tiled: null pointer");
    CNA_REQUIRE(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) |
                  reinterpret_cast<uintptr_t>(usrc)) & 15) == 0,
                "cna_diffuse_step_f32_tiled: state and source lists must be 16-byte aligned");
    if (n_rows == 0 || n_tiles == 0) return CNA_OK;
    Args a{};
    a.indptr = indptr;
    a.epair = static_cast<const uint2 *>(epair);
    a.diag = diag;
    a.in = in;
    a.out = out;
    a.ld4 = ld / 4;
    a.nvec = (n_cols + 3) / 4;
    a.n_slabs = (a.nvec + kSlabVec - 1) / kSlabVec;
    a.in_row_offset = in_row_offset;
    a.tile_row = tile_row;
    a.tile_u = tile_u;
    a.usrc = usrc;
    a.n_tiles = n_tiles;
    a.stage_mode = stage_mode;
    CUtensorMap tm;
    {
        EncodeTiledFn fn = encode_fn();
        if (!fn) return set_error(CNA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        static const int box_rows = getenv("CNA_TILED_BOX_ROWS") ? atoi(getenv("CNA_TILED_BOX_ROWS")) : 1;
        cuuint64_t dims[2] = {cuuint64_t(ld), cuuint64_t(in_rows)};
        cuuint64_t strides[1] = {cuuint64_t(ld) * 4};
        cuuint32_t box[2] = {cuuint32_t(kSlabVec * 4), cuuint32_t(box_rows)};
        cuuint32_t estr[2] = {1, 1};
        CUresult rc = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(in), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) return set_error(CNA_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", int(rc));
    }
    const size_t smem = size_t(2) * kBufBytes + 128 + 64;
    CNA_CUDA(cudaFuncSetAttribute(spmm_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    const unsigned grid = unsigned(n_tiles < num_sms() ? n_tiles : num_sms());
    spmm_tiled_kernel<<<grid, kThreads, smem, as_stream(stream)>>>(tm, a);
    CNA_LAUNCHED("spmm_tiled_kernel");
    return CNA_OK;
}

}  // extern "C"
