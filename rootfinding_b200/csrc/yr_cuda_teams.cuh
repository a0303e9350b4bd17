// CUDA teams of the frontier pipelines (yr_f2.h, yr_fd.h): a team of 16 or 32 lanes of one warp, or a whole CTA, behind the
// small interface the kernel source is written against (tid, nthreads, sync, sums, shuffles, asynchronous copies).
#pragma once
#include <cuda_runtime.h>
#include <cuda_pipeline.h>
#include <stdint.h>

namespace yr_cuda {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// asynchronous global -> shared copy of n_even doubles (both 16-byte aligned): LDGSTS, no registers in between
__device__ __forceinline__ void copy_async16(double* dst, const double* src, int n_even, int tid, int nthreads) {
    const int n2 = n_even >> 1;
    for (int e = tid; e < n2; e += nthreads) __pipeline_memcpy_async(dst + 2 * e, src + 2 * e, 16);
    __pipeline_commit();
}

// ---- TMA bulk copies (cp.async.bulk, sm_90+): one lane moves a whole tensor between an arena and shared memory ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Measured on the bench workload (profiles/r02_experiments.txt): bulk STORES of finished tensors are a small gain
// (one issuing lane instead of a copy loop), bulk LOADS of these 1-3 KB tensors are a loss against LDGSTS (the team
// waits on the mbarrier for longer than the per-lane copies take).
#ifndef YR_F2_TMA_LOAD
#define YR_F2_TMA_LOAD 0
#endif
#ifndef YR_F2_TMA_STORE
#define YR_F2_TMA_STORE 1
#endif

// A team of LANES (32 or 16) lanes of one warp.  Two 16-lane teams share a warp: they run the same instruction
// stream on different boxes (every shuffle / vote / barrier names only the team's own lanes), which halves the
// per-box instruction cost of everything that cannot fill 32 lanes anyway (small boxes).
template <int LANES>
struct WarpTeam {
    static constexpr bool kShuffle = true;
    int tid, nthreads, lane, lanes, warp, warps;
    unsigned mask;
    uint64_t* bar;            // the team's mbarrier (TMA loads)
    uint32_t phase;           // its parity
    bool tma_pending, store_pending;
    __device__ __forceinline__ void sync() { __syncwarp(mask); }
    __device__ __forceinline__ void wsync() { __syncwarp(mask); }
    __device__ __forceinline__ double wsum(double v) {
#pragma unroll
        for (int off = LANES / 2; off > 0; off >>= 1) v += __shfl_xor_sync(mask, v, off);
        return v;
    }
    __device__ __forceinline__ double sum(double v) { return wsum(v); }
    __device__ __forceinline__ double sum_sync(double v) { __syncwarp(mask); return wsum(v); }
    __device__ __forceinline__ void sum2_sync(double a, double b, double& ra, double& rb) { __syncwarp(mask); ra = wsum(a); rb = wsum(b); }
    __device__ __forceinline__ bool wany(bool p) { return (__ballot_sync(mask, p) & mask) != 0u; }
    __device__ __forceinline__ unsigned wballot(bool p) { return LANES == 32 ? __ballot_sync(mask, p) : ((__ballot_sync(mask, p) >> (threadIdx.x & 16)) & 0xffffu); }
    __device__ __forceinline__ double wsum_w(double v, int w) { for (int off = w >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(mask, v, off); return v; }
    __device__ __forceinline__ double wshfl_idx_d(double v, int src) { return __shfl_sync(mask, v, src, LANES); }
    __device__ __forceinline__ double wshfl_up(double v, int d, int w) { return __shfl_up_sync(mask, v, d, w < LANES ? w : LANES); }
    __device__ __forceinline__ double wshfl_down(double v, int d, int w) { return __shfl_down_sync(mask, v, d, w < LANES ? w : LANES); }
    __device__ __forceinline__ int wshfl_idx(int v, int src, int w) { return __shfl_sync(mask, v, src, w < LANES ? w : LANES); }
    __device__ __forceinline__ unsigned long long bcast(unsigned long long v) { return __shfl_sync(mask, v, 0, LANES); }
    __device__ __forceinline__ unsigned long long atomic_add(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
    // global -> shared, both 16-byte aligned, n_even even.  The team must have synchronised after its last use of dst.
    __device__ __forceinline__ void copy_async(double* dst, const double* src, int n_even) {
        if (YR_F2_TMA_LOAD && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0 && n_even > 0) {
            if (tid == 0) {
                fence_proxy_async();                      // earlier generic-proxy accesses to dst are ordered before the bulk write
                mbar_arrive_expect_tx(bar, (uint32_t)n_even * 8u);
                tma_bulk_g2s(dst, src, (uint32_t)n_even * 8u, bar);
            }
            tma_pending = true;
        } else copy_async16(dst, src, n_even, tid, LANES);
    }
    __device__ __forceinline__ void copy_wait() {
        if (tma_pending) { while (!mbar_try_wait(bar, phase)) {} phase ^= 1u; tma_pending = false; }
        __pipeline_wait_prior(0);
    }
    // shared -> global of a flat block; the caller's data must be complete in shared memory (this synchronises the team)
    __device__ __forceinline__ void store_async(double* dst, const double* src, int n_even) {
        if (YR_F2_TMA_STORE && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0 && n_even > 0) {
            __syncwarp(mask);
            if (tid == 0) { fence_proxy_async(); tma_bulk_s2g(dst, src, (uint32_t)n_even * 8u); }
            store_pending = true;
        } else {
            const double2* s2 = reinterpret_cast<const double2*>(src); double2* d2 = reinterpret_cast<double2*>(dst);
            if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) for (int e = tid; e < (n_even >> 1); e += LANES) d2[e] = s2[e];
            else for (int e = tid; e < n_even; e += LANES) dst[e] = src[e];
        }
    }
    // the shared-memory sources of every store_async so far may be overwritten after this
    __device__ __forceinline__ void store_wait() {
        if (store_pending) { if (tid == 0) tma_store_wait_read(); store_pending = false; }
    }
    __device__ __forceinline__ void store_drain() { if (tid == 0) tma_store_wait_all(); }
};
template <int LANES>
__device__ __forceinline__ WarpTeam<LANES> make_warp_team(uint64_t* bars) {
    WarpTeam<LANES> tm;
    tm.tid = tm.lane = threadIdx.x & (LANES - 1); tm.nthreads = tm.lanes = LANES; tm.warp = 0; tm.warps = 1;
    tm.mask = LANES == 32 ? 0xffffffffu : (0xffffu << (threadIdx.x & 16));
    tm.bar = bars + threadIdx.x / LANES; tm.phase = 0; tm.tma_pending = false; tm.store_pending = false;
    if (tm.tid == 0) mbar_init(tm.bar, 1);
    fence_mbar_init();
    __syncthreads();
    return tm;
}

struct CtaShared { double red[2][16]; unsigned long long slot; };

struct CtaTeam {
    static constexpr bool kShuffle = true;
    int tid, nthreads, lane, lanes, warp, warps;
    int flip;
    CtaShared* sh;
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ void wsync() { __syncwarp(); }
    // one barrier per reduction: partials alternate between two buffers, so a buffer is rewritten only after
    // every thread has passed the barrier of the reduction in between
    __device__ __forceinline__ double sum_sync(double v) {
        v = warp_sum_d(v);
        const int buf = flip; flip ^= 1;
        if (lane == 0) sh->red[buf][warp] = v;
        __syncthreads();
        double t = 0.0;
        for (int w = 0; w < warps; ++w) t += sh->red[buf][w];
        return t;
    }
    __device__ __forceinline__ void sum2_sync(double a, double b, double& ra, double& rb) {
        a = warp_sum_d(a); b = warp_sum_d(b);
        const int buf = flip; flip ^= 1;
        if (lane == 0) { sh->red[buf][warp] = a; sh->red[buf][8 + warp] = b; }
        __syncthreads();
        double ta = 0.0, tb = 0.0;
        for (int w = 0; w < warps; ++w) { ta += sh->red[buf][w]; tb += sh->red[buf][8 + w]; }
        ra = ta; rb = tb;
    }
    __device__ __forceinline__ double sum(double v) { return sum_sync(v); }
    __device__ __forceinline__ double wsum(double v) { return warp_sum_d(v); }
    __device__ __forceinline__ bool wany(bool p) { return __any_sync(0xffffffffu, p); }
    __device__ __forceinline__ unsigned wballot(bool p) { return __ballot_sync(0xffffffffu, p); }
    __device__ __forceinline__ double wsum_w(double v, int w) { for (int off = w >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off); return v; }
    __device__ __forceinline__ double wshfl_idx_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
    __device__ __forceinline__ double wshfl_up(double v, int d, int w) { return __shfl_up_sync(0xffffffffu, v, d, w); }
    __device__ __forceinline__ double wshfl_down(double v, int d, int w) { return __shfl_down_sync(0xffffffffu, v, d, w); }
    __device__ __forceinline__ int wshfl_idx(int v, int src, int w) { return __shfl_sync(0xffffffffu, v, src, w); }
    __device__ __forceinline__ unsigned long long bcast(unsigned long long v) {
        if (tid == 0) sh->slot = v;
        __syncthreads();
        const unsigned long long r = sh->slot;
        __syncthreads();
        return r;
    }
    __device__ __forceinline__ unsigned long long atomic_add(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
    __device__ __forceinline__ void copy_async(double* dst, const double* src, int n_even) { copy_async16(dst, src, n_even, tid, nthreads); }
    __device__ __forceinline__ void copy_wait() { __pipeline_wait_prior(0); }
    __device__ __forceinline__ void store_async(double* dst, const double* src, int n_even) { for (int e = tid; e < n_even; e += nthreads) dst[e] = src[e]; }
    __device__ __forceinline__ void store_wait() {}
};
__device__ __forceinline__ CtaTeam make_cta_team(CtaShared* sh) {
    CtaTeam tm; tm.tid = threadIdx.x; tm.nthreads = blockDim.x; tm.lane = threadIdx.x & 31; tm.lanes = 32;
    tm.warp = threadIdx.x >> 5; tm.warps = blockDim.x >> 5; tm.flip = 0; tm.sh = sh; return tm;
}


struct DevAt {
    __device__ __forceinline__ unsigned long long add(unsigned long long* p, unsigned long long v) const { return atomicAdd(p, v); }
    __device__ __forceinline__ void max(unsigned long long* p, unsigned long long v) const { atomicMax(p, v); }
};

__device__ __forceinline__ unsigned long long warp_add(unsigned long long v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ unsigned long long warp_max(unsigned long long v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, off); v = o > v ? o : v; }
    return v;
}

}  // namespace yr_cuda
