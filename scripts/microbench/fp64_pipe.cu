// fp64 pipe on B200: dependent-issue latency (1 warp) and throughput (32 warps / SM) of DFMA, DADD, DSETP,
// and an LDS.64 -> DADD chain.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe fp64_pipe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double *out, long long *clk, int iters, double seed) {
    __shared__ double sm[1024];
    sm[threadIdx.x] = seed + threadIdx.x;
    __syncthreads();
    double a = seed + threadIdx.x, b = 1.0000001, c = 1e-9;
    int cnt = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (MODE == 0) a = fma(a, b, c);
            if (MODE == 1) a = a + c;
            if (MODE == 2) { cnt += (a < b) ? 1 : 0; b += 1.0; }             // DSETP + DADD (independent-ish)
            if (MODE == 3) a = a + sm[(threadIdx.x + u * 32 + i) & 1023];   // LDS -> DADD
            if (MODE == 4) a = a * b;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + cnt + b;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
    double *out; long long *clk, h;
    cudaMalloc(&out, 8 * 1024 * 256); cudaMalloc(&clk, 8);
    const char *names[] = {"DFMA chain", "DADD chain", "DSETP+DADD", "LDS->DADD chain", "DMUL chain"};
    for (int mode = 0; mode < 5; ++mode)
        for (int threads : {32, 128, 512, 1024}) {
            const int iters = 2000;
            switch (mode) {
                case 0: k<0><<<1, threads>>>(out, clk, iters, 1.0); break;
                case 1: k<1><<<1, threads>>>(out, clk, iters, 1.0); break;
                case 2: k<2><<<1, threads>>>(out, clk, iters, 1.0); break;
                case 3: k<3><<<1, threads>>>(out, clk, iters, 1.0); break;
                case 4: k<4><<<1, threads>>>(out, clk, iters, 1.0); break;
            }
            cudaDeviceSynchronize();
            cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
            printf("%-16s %4d threads: %.2f clk per op per warp (%.2f lane-ops/clk/SM)\n", names[mode], threads,
                   double(h) / (iters * 16.0), threads * iters * 16.0 / double(h));
        }
    return 0;
}
