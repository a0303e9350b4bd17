// Microbenchmark: FP64 vector throughput on one B200 -- DFMA, separately rounded DMUL+DADD, and the
// shared-memory-fed MAC loop shape the restriction kernels use.  Prints GFLOP/s (1 FMA = 2 flop,
// 1 MUL + 1 ADD = 2 flop).  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k_reg(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
            x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
            x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
        } else {
            x0 = __dadd_rn(__dmul_rn(x0, a), b); x1 = __dadd_rn(__dmul_rn(x1, a), b); x2 = __dadd_rn(__dmul_rn(x2, a), b);
            x3 = __dadd_rn(__dmul_rn(x3, a), b); x4 = __dadd_rn(__dmul_rn(x4, a), b); x5 = __dadd_rn(__dmul_rn(x5, a), b);
            x6 = __dadd_rn(__dmul_rn(x6, a), b); x7 = __dadd_rn(__dmul_rn(x7, a), b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// smem-fed: acc_r += C[r][k] * M[k][lane]  for R accumulators per lane (R-row blocking), separate mul/add
template <int R, bool FMA>
__global__ void __launch_bounds__(256) k_smem(double* out, int iters, int n) {
    extern __shared__ double sm[];
    double* C = sm;                 // R * n (broadcast reads)
    double* M = sm + 8 * 64;        // n * 32 per warp
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 64 + 2 * 64 * 32; i += blockDim.x) sm[i] = 1.0 / (1 + i);
    __syncthreads();
    const double* Mw = M + (w & 1) * n * 32;
    double acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
        for (int k = 0; k < n; ++k) {
            const double x = Mw[k * 32 + lane];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (FMA) acc[r] = __fma_rn(C[r * 64 + k], x, acc[r]);
                else acc[r] = __dadd_rn(acc[r], __dmul_rn(C[r * 64 + k], x));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) s += acc[r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_it(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
    double* out; cudaMalloc(&out, 148 * 16 * 256 * 8);
    const int grid = 148 * 8, iters = 20000;
    float ms = time_it([&] { k_reg<0><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
    printf("DFMA           : %.1f GFLOP/s\n", 2.0 * 8 * iters * grid * 256 / ms * 1e-6);
    ms = time_it([&] { k_reg<1><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
    printf("DMUL+DADD      : %.1f GFLOP/s\n", 2.0 * 8 * iters * grid * 256 / ms * 1e-6);
    const int n = 64, it2 = 400;
    const size_t smem = (8 * 64 + 2 * 64 * 32) * 8;
#define RUN(R, FMA)                                                                                         \
    cudaFuncSetAttribute(k_smem<R, FMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
    ms = time_it([&] { k_smem<R, FMA><<<148 * 6, 256, smem>>>(out, it2, n); });                                 \
    printf("smem MAC R=%d %s: %.1f GFLOP/s (6 CTAs of 8 warps per SM)\n", R, FMA ? "fma    " : "mul+add", 2.0 * R * n * it2 * 148 * 6 * 256 / ms * 1e-6);
    RUN(1, false) RUN(2, false) RUN(4, false) RUN(8, false) RUN(4, true) RUN(8, true)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
