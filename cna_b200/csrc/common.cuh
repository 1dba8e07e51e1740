// Shared helpers for the cna_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/cna_b200.h"

namespace cna {

int set_error(int code, const char *fmt, ...);
void count_launch(int n = 1);

#define CNA_REQUIRE(cond, ...)                                        \
    do {                                                              \
        if (!(cond)) return cna::set_error(CNA_ERR_INVALID, __VA_ARGS__); \
    } while (0)

#define CNA_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return cna::set_error(CNA_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                                  __FILE__, __LINE__);                                      \
    } while (0)

#define CNA_LAUNCHED(name)                                                                   \
    do {                                                                                     \
        cudaError_t e_ = cudaGetLastError();                                                 \
        if (e_ != cudaSuccess)                                                               \
            return cna::set_error(CNA_ERR_CUDA, "launch of %s failed: %s", name,             \
                                  cudaGetErrorString(e_));                                   \
        cna::count_launch();                                                                 \
    } while (0)

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// scipy.stats.kurtosis (biased) from central moments: NaN when the variance is negligible
// relative to the mean (scipy/stats/_stats_py.py: `zero = m2 <= (eps * mean)**2`).
__device__ __forceinline__ double kurtosis_from_moments(double mean, double m2, double m4, bool fisher) {
    const double eps = 2.220446049250313e-16;
    double lim = eps * mean;
    if (m2 <= lim * lim) return nan("");
    double k = m4 / (m2 * m2);
    return fisher ? k - 3.0 : k;
}

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

inline int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

}  // namespace cna
