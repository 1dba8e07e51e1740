// tcgen05 (5th-generation tensor core) versions of the Gram and the null-correlation GEMM.
// Placeholder translation unit: the SIMT kernels in gemm_simt.cu are used until these land.
#include "common.cuh"

namespace cna {

bool tc_enabled() { return false; }

int gram_tc(const float *, int64_t, int64_t, int, double *, cudaStream_t) { return -1; }

int null_hist_tc(const float *, int64_t, int64_t, int, const float *, int64_t, int, const double *, int,
                 uint32_t *, cudaStream_t) {
    return -1;
}

}  // namespace cna
