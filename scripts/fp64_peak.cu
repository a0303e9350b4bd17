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

// FP64 tensor-core path: mma.sync m8n8k4 (SASS DMMA.8x8x4), register-resident operands, NACC independent accumulators.
// One instruction = 8 x 8 x 4 fused multiply-adds = 512 flop per warp.
template <int NACC>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a0, double b0) {
    double a = a0 + threadIdx.x * 1e-9, b = b0 - threadIdx.x * 1e-9;
    double c[NACC][2];
#pragma unroll
    for (int q = 0; q < NACC; ++q) { c[q][0] = q; c[q][1] = -q; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < NACC; ++q)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[q][0]), "+d"(c[q][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < NACC; ++q) s += c[q][0] + c[q][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// The subdivision contraction as DMMA tiles: out[8 rows][8 fibers] += C[8 rows][4 k] * M[4 k][8 fibers], C (the fixed
// level-independent matrix) and M (a box's tensor) both in shared memory, k swept over n -- what a tensor-core version
// of the mirrored subdivide sweep would execute per (8-row block, 8-fiber group), fused rounding (not bit-identical).
__global__ void __launch_bounds__(256) k_dmma_smem(double* out, int iters, int n) {
    extern __shared__ double sm[];
    double* C = sm;                 // 64 x 64 matrix, row-major
    double* M = sm + 64 * 64;       // n x 64 tensor per CTA (8 warps x 8 fibers)
    for (int i = threadIdx.x; i < 64 * 64 + 64 * 64; i += blockDim.x) sm[i] = 1.0 / (1 + i);
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;               // fragment coordinates of m8n8k4: a = A[g][t], b = B[t][g]
    double acc[8][2];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q][0] = acc[q][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
        for (int k = 0; k < n; k += 4) {
            const double b = M[(k + t) * 64 + w * 8 + g];
#pragma unroll
            for (int q = 0; q < 8; ++q) {                // 8 row blocks of 8 rows share the M fragment
                const double a = C[(q * 8 + g) * 64 + k + t];
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(acc[q][0]), "+d"(acc[q][1]) : "d"(a), "d"(b));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += acc[q][0] + acc[q][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// dependent chains: cycles per DADD / DMUL / DFMA when every instruction waits for the previous one (one warp per SM)
template <int MODE>
__global__ void k_chain(double* out, long long* cycles, int iters, double a, double b) {
    double x = threadIdx.x;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) x = __dadd_rn(x, b);
        else if (MODE == 1) x = __dmul_rn(x, a);
        else x = __fma_rn(x, a, b);
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
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
    // the same R = 4 loop at the occupancies the frontier kernels run at (1, 2, 4 CTAs of 4 warps per SM)
    for (int ctas = 1; ctas <= 4; ctas *= 2) {
        cudaFuncSetAttribute(k_smem<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ms = time_it([&] { k_smem<4, false><<<148 * ctas, 128, smem>>>(out, it2, n); });
        printf("smem MAC R=4 mul+add, %2d warps per SM: %.1f GFLOP/s\n", 4 * ctas, 2.0 * 4 * n * it2 * 148 * ctas * 128 / ms * 1e-6);
    }
    // FP64 tensor cores
    ms = time_it([&] { k_dmma<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
    printf("DMMA.8x8x4 (8 accumulators, registers): %.1f GFLOP/s\n", 512.0 * 8 * iters * grid * 8 / ms * 1e-6);
    ms = time_it([&] { k_dmma<2><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
    printf("DMMA.8x8x4 (2 accumulators, registers): %.1f GFLOP/s\n", 512.0 * 2 * iters * grid * 8 / ms * 1e-6);
    {
        const size_t sm2 = (64 * 64 + 64 * 64) * 8;
        cudaFuncSetAttribute(k_dmma_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
        const int it3 = 2000;
        ms = time_it([&] { k_dmma_smem<<<148 * 3, 256, sm2>>>(out, it3, 64); });
        printf("DMMA contraction from shared memory (64 x 64 matrix x 64 fibers, 3 CTAs of 8 warps per SM): %.1f GFLOP/s\n",
               512.0 * 8 * (64 / 4) * it3 * 148.0 * 3 * 8 / ms * 1e-6);
    }
    // dependent-chain latencies
    {
        long long* cyc; cudaMalloc(&cyc, 8);
        const char* names[3] = {"DADD", "DMUL", "DFMA"};
        long long h;
        const int itc = 100000;
        k_chain<0><<<1, 32>>>(out, cyc, itc, 1.0000001, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("dependent %s chain: %.1f cycles per instruction\n", names[0], (double)h / itc);
        k_chain<1><<<1, 32>>>(out, cyc, itc, 1.0000001, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("dependent %s chain: %.1f cycles per instruction\n", names[1], (double)h / itc);
        k_chain<2><<<1, 32>>>(out, cyc, itc, 1.0000001, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("dependent %s chain: %.1f cycles per instruction\n", names[2], (double)h / itc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
