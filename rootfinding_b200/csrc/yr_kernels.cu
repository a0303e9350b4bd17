// CUDA backend (sm_100a) of include/yroots_b200.h.
//
// Kernels are persistent: a launch has at most (SM count x resident CTAs per SM) CTAs, each
// pulling boxes from an atomic cursor, so grids are sized to the 148 SMs of a B200 rather than
// to the number of boxes.  A box's tensors are contiguous in the level pool and are staged
// into shared memory with TMA bulk copies (cp.async.bulk + mbarrier complete_tx); all FP64
// work on them (reductions, restriction-matrix recurrences, triangular contractions) then runs
// out of shared memory, and only the children go back to HBM.  Compiled with -fmad=false:
// products and sums are rounded separately (like the reference's numba code) except where
// fma() is written explicitly.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "yr_api_impl.h"
#include "yr_approx.h"

namespace {

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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "YR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra YR_DONE;\n"
        "bra YR_WAIT;\n"
        "YR_DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

struct DeviceCtx {
    static constexpr bool kWarpPerBox = false;
    int tid, nthreads, lane, lanes, warp, nwarps;
    double* red;          // [32] shared scratch for block reductions
    uint64_t* mbar;       // shared mbarrier for TMA staging
    uint32_t phase;       // its current parity (uniform across the CTA)
    int ws_is_smem;

    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ double warp_tree(double v) {
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        return v;
    }
    __device__ __forceinline__ double block_sum(double v) {
        v = warp_tree(v);
        if (lane == 0) red[warp] = v;
        __syncthreads();
        double t = 0.0;
        for (int w = 0; w < nwarps; ++w) t += red[w];
        __syncthreads();
        return t;
    }
    __device__ __forceinline__ double block_max(double v) {
        for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, off));
        if (lane == 0) red[warp] = v;
        __syncthreads();
        double t = red[0];
        for (int w = 1; w < nwarps; ++w) t = fmax(t, red[w]);
        __syncthreads();
        return t;
    }
    // maximum of non-negative values that keeps a NaN (numpy's max propagates it)
    __device__ __forceinline__ double block_max_nan(double v) {
        for (int off = 16; off > 0; off >>= 1) { const double o = __shfl_down_sync(0xffffffffu, v, off); v = (v != v || o != o) ? NAN : fmax(v, o); }
        if (lane == 0) red[warp] = v;
        __syncthreads();
        double t = red[0];
        for (int w = 1; w < nwarps; ++w) t = (t != t || red[w] != red[w]) ? NAN : fmax(t, red[w]);
        __syncthreads();
        return t;
    }
    __device__ __forceinline__ double warp_sum(double v) {
        v = warp_tree(v);
        return __shfl_sync(0xffffffffu, v, 0);
    }
    __device__ __forceinline__ unsigned long long atomic_add(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
    __device__ __forceinline__ void atomic_max(unsigned long long* p, unsigned long long v) { atomicMax(p, v); }

    // Stage n doubles (global -> workspace).  Collective.  With the workspace in shared memory and
    // 16-byte aligned endpoints the copy is a TMA bulk transfer tracked by an mbarrier.
    __device__ __forceinline__ void copy_in(double* dst, const double* src, long long n) {
        const bool tma_ok = ws_is_smem && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                            ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && n >= 2;
        if (tma_ok) {
            const long long nbulk = n & ~1ll;                    // multiple of 16 bytes
            if (tid == 0) {
                const unsigned long long bytes = (unsigned long long)nbulk * 8ull;
                fence_proxy_async();       // earlier generic-proxy accesses to this workspace are ordered before the bulk write
                mbar_arrive_expect_tx(mbar, (uint32_t)bytes);
                unsigned long long done = 0;
                while (done < bytes) {
                    const uint32_t chunk = (uint32_t)((bytes - done) > 65536ull ? 65536ull : (bytes - done));
                    tma_bulk_g2s(reinterpret_cast<char*>(dst) + done, reinterpret_cast<const char*>(src) + done, chunk, mbar);
                    done += chunk;
                }
            }
            if (tid == 32 % nthreads && (n & 1)) dst[n - 1] = src[n - 1];
            mbar_wait(mbar, phase);
            phase ^= 1u;
        } else {
            for (long long i = tid; i < n; i += nthreads) dst[i] = src[i];
        }
    }
    __device__ __forceinline__ double cospi_ratio(long long m, long long N) const { return cospi((double)m / (double)N); }
};

__device__ __forceinline__ DeviceCtx make_ctx(double* red, uint64_t* mbar, int ws_is_smem) {
    DeviceCtx cx;
    cx.tid = threadIdx.x; cx.nthreads = blockDim.x; cx.lane = threadIdx.x & 31; cx.lanes = 32;
    cx.warp = threadIdx.x >> 5; cx.nwarps = blockDim.x >> 5;
    cx.red = red; cx.mbar = mbar; cx.phase = 0; cx.ws_is_smem = ws_is_smem;
    if (threadIdx.x == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    __syncthreads();
    return cx;
}

__host__ __device__ constexpr size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// Warp-per-box variant of the execution context: one warp is a whole "CTA" of the pipeline.
// Barriers become __syncwarp, block reductions one shuffle tree; a serial section (BILS, checks)
// stalls only its own warp while the other warps of the SM work on their own boxes.
struct WarpCtx {
    static constexpr bool kWarpPerBox = true;
    int tid, nthreads, lane, lanes, warp, nwarps;
    uint64_t* mbar;
    uint32_t phase;

    __device__ __forceinline__ void sync() { __syncwarp(); }
    __device__ __forceinline__ void cta_sync() { __syncthreads(); }
    __device__ __forceinline__ bool cta_all(bool pred) { return __syncthreads_and(pred ? 1 : 0) != 0; }
    __device__ __forceinline__ double block_sum(double v) {
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        return __shfl_sync(0xffffffffu, v, 0);
    }
    __device__ __forceinline__ double block_max(double v) {
        for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, off));
        return __shfl_sync(0xffffffffu, v, 0);
    }
    __device__ __forceinline__ double warp_sum(double v) { return block_sum(v); }
    __device__ __forceinline__ double shfl_up(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
    __device__ __forceinline__ double shfl_down(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
    __device__ __forceinline__ int shfl_idx_int(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
    __device__ __forceinline__ unsigned long long atomic_add(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
    __device__ __forceinline__ void atomic_max(unsigned long long* p, unsigned long long v) { atomicMax(p, v); }
    __device__ __forceinline__ void copy_in(double* dst, const double* src, long long n) {
        const bool tma_ok = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && n >= 2;
        if (tma_ok) {
            const long long nbulk = n & ~1ll;
            if (lane == 0) {
                const unsigned long long bytes = (unsigned long long)nbulk * 8ull;
                fence_proxy_async();
                mbar_arrive_expect_tx(mbar, (uint32_t)bytes);
                unsigned long long done = 0;
                while (done < bytes) {
                    const uint32_t chunk = (uint32_t)((bytes - done) > 65536ull ? 65536ull : (bytes - done));
                    tma_bulk_g2s(reinterpret_cast<char*>(dst) + done, reinterpret_cast<const char*>(src) + done, chunk, mbar);
                    done += chunk;
                }
            }
            if (lane == 1 && (n & 1)) dst[n - 1] = src[n - 1];
            mbar_wait(mbar, phase);
            phase ^= 1u;
        } else {
            for (long long i = lane; i < n; i += 32) dst[i] = src[i];
        }
    }
    __device__ __forceinline__ double cospi_ratio(long long m, long long N) const { return cospi((double)m / (double)N); }
};

template <int D>
__global__ void __launch_bounds__(256) init_kernel(const yr::InitArgs<D> args, int ws_in_smem) {
    extern __shared__ __align__(16) double dyn_ws[];
    __shared__ yr::BoxShared<D> sh;
    __shared__ double red[32];
    __shared__ __align__(8) uint64_t mbar;
    DeviceCtx cx = make_ctx(red, &mbar, ws_in_smem);
    double* ws = ws_in_smem ? dyn_ws : args.gws + (size_t)blockIdx.x * args.gws_stride;
    yr::init_worker<D>(cx, args, sh, ws);
}

// ---- phase-split pipeline kernels (yr_phases.h) ---------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) tensor_step_kernel(const yr::PhaseArgs<D> pa, int ws_in_smem) {
    extern __shared__ __align__(16) double dyn_ws[];
    __shared__ yr::BoxShared<D> sh;
    __shared__ double red[32];
    __shared__ __align__(8) uint64_t mbar;
    DeviceCtx cx = make_ctx(red, &mbar, ws_in_smem);
    double* ws = ws_in_smem ? dyn_ws : pa.lv.gws + (size_t)blockIdx.x * pa.lv.gws_stride;
    yr::tensor_step_worker<D>(cx, pa, sh, ws);
}

template <int D>
__global__ void __launch_bounds__(256) subdivide_step_kernel(const yr::PhaseArgs<D> pa, int ws_in_smem) {
    extern __shared__ __align__(16) double dyn_ws[];
    __shared__ yr::BoxShared<D> sh;
    __shared__ double red[32];
    __shared__ __align__(8) uint64_t mbar;
    DeviceCtx cx = make_ctx(red, &mbar, ws_in_smem);
    double* ws = ws_in_smem ? dyn_ws : pa.lv.gws + (size_t)blockIdx.x * pa.lv.gws_stride;
    yr::subdivide_step_worker<D>(cx, pa, sh, ws);
}

template <int D>
__device__ __forceinline__ WarpCtx make_warp_ctx(unsigned char* dyn_raw, unsigned warp_stride_bytes, yr::BoxShared<D>*& sh, double*& ws) {
    const int w = threadIdx.x >> 5;
    unsigned char* base = dyn_raw + (size_t)w * warp_stride_bytes;
    sh = reinterpret_cast<yr::BoxShared<D>*>(base);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(base + align16(sizeof(yr::BoxShared<D>)));
    ws = reinterpret_cast<double*>(base + align16(sizeof(yr::BoxShared<D>)) + 16);
    WarpCtx cx;
    cx.tid = cx.lane = threadIdx.x & 31; cx.nthreads = cx.lanes = 32; cx.warp = 0; cx.nwarps = 1;
    cx.mbar = mbar; cx.phase = 0;
    if (cx.lane == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    __syncwarp();
    return cx;
}

template <int D>
__global__ void __launch_bounds__(128, 4) tensor_step_kernel_warp(const yr::PhaseArgs<D> pa, unsigned warp_stride_bytes) {
    extern __shared__ __align__(16) unsigned char dyn_raw[];
    yr::BoxShared<D>* sh; double* ws;
    WarpCtx cx = make_warp_ctx<D>(dyn_raw, warp_stride_bytes, sh, ws);
    yr::tensor_step_worker<D>(cx, pa, *sh, ws);
}

template <int D>
__global__ void __launch_bounds__(128, 4) subdivide_step_kernel_warp(const yr::PhaseArgs<D> pa, unsigned warp_stride_bytes) {
    extern __shared__ __align__(16) unsigned char dyn_raw[];
    yr::BoxShared<D>* sh; double* ws;
    WarpCtx cx = make_warp_ctx<D>(dyn_raw, warp_stride_bytes, sh, ws);
    yr::subdivide_step_worker<D>(cx, pa, *sh, ws);
}

struct DeviceAtomics {
    __device__ __forceinline__ unsigned long long add(unsigned long long* p, unsigned long long v) const { return atomicAdd(p, v); }
};

// thread per box: the serial per-box mathematics, 32 boxes per instruction stream
template <int D>
__global__ void __launch_bounds__(128) scalar_step_kernel(const yr::PhaseArgs<D> pa) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    yr::ScalarCounts c{};
    if (i < pa.n_list) yr::scalar_step<D>(pa, pa.list_in ? pa.list_in[i] : (unsigned int)i, c, DeviceAtomics());
    unsigned long long v[6] = {c.boxes, c.zooms, c.dconst, c.dquad, c.dzoom, c.coeffs};
    unsigned long long mx = c.maxlevel;
#pragma unroll
    for (int k = 0; k < 6; ++k)
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], off);
    for (int off = 16; off > 0; off >>= 1) { const unsigned long long o = __shfl_down_sync(0xffffffffu, mx, off); mx = o > mx ? o : mx; }
    if ((threadIdx.x & 31) == 0) {
        yr::LevelCounters* k = pa.lv.ctr;
        if (v[0]) atomicAdd(&k->boxes, v[0]);
        if (v[1]) atomicAdd(&k->zooms, v[1]);
        if (v[2]) atomicAdd(&k->discard_const, v[2]);
        if (v[3]) atomicAdd(&k->discard_quad, v[3]);
        if (v[4]) atomicAdd(&k->discard_zoom, v[4]);
        if (v[5]) atomicAdd(&k->entry_coeffs, v[5]);
        if (mx) atomicMax(&k->max_level, mx);
    }
    unsigned long long m6[6] = {c.t_total, c.t_tensor, c.t_extent, c.s_total, c.s_tensor, c.s_extent};
#pragma unroll
    for (int k = 0; k < 6; ++k)
        for (int off = 16; off > 0; off >>= 1) { const unsigned long long o = __shfl_down_sync(0xffffffffu, m6[k], off); m6[k] = o > m6[k] ? o : m6[k]; }
    if ((threadIdx.x & 31) == 0) {
        yr::RoundCtl* r = pa.ctl;
        if (m6[0]) { atomicMax(&r->max_total, m6[0]); atomicMax(&r->max_tensor, m6[1]); atomicMax(&r->max_extent, m6[2]); }
        if (m6[3]) { atomicMax(&r->sub_max_total, m6[3]); atomicMax(&r->sub_max_tensor, m6[4]); atomicMax(&r->sub_max_extent, m6[5]); }
    }
}

template <int D>
__global__ void bils_kernel(long long n, const double* lin, const double* c0, const double* sumabs, const double* errs,
                            const int* final_step, double* interval, int* flags) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) yr::bils_item<D>(i, lin, c0, sumabs, errs, final_step, interval, flags);
}

template <int D>
__global__ void quad_kernel(long long n, const double* coeffs, const long long* off, const int* shapes, const double* tol,
                            int nd_check, int* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) yr::quad_item<D>(i, coeffs, off, shapes, tol, nd_check, out);
}

__global__ void dct1_kernel(yr::Dct1Args a) {
    extern __shared__ __align__(16) double table[];
    __shared__ double red[32];
    __shared__ __align__(8) uint64_t mbar;
    DeviceCtx cx = make_ctx(red, &mbar, 0);
    yr::dct1_worker(cx, a, table, blockIdx.x, gridDim.x);
}

__global__ void dct1_table_kernel(double* table, long long N) {
    for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < 2 * N; m += (long long)gridDim.x * blockDim.x)
        table[m] = cospi((double)m / (double)N);
}
__global__ void dct1_kernel_gtable(yr::Dct1Args a, double* table) {
    __shared__ double red[32];
    __shared__ __align__(8) uint64_t mbar;
    DeviceCtx cx = make_ctx(red, &mbar, 0);
    yr::dct1_worker(cx, a, table, blockIdx.x, gridDim.x, false);
}

__global__ void poly_eval_kernel(yr::PolyEvalArgs a) {
    __shared__ double red[32];
    __shared__ __align__(8) uint64_t mbar;
    DeviceCtx cx = make_ctx(red, &mbar, 0);
    yr::poly_eval_worker(cx, a, blockIdx.x, gridDim.x);
}

__global__ void upper_axis_kernel(yr::UpperAxisArgs a) {
    __shared__ double red[32];
    __shared__ __align__(8) uint64_t mbar;
    DeviceCtx cx = make_ctx(red, &mbar, 0);
    yr::upper_axis_worker(cx, a, blockIdx.x, gridDim.x);
}

__global__ void abs_mean_kernel(yr::AbsMeanArgs a) {
    __shared__ double red[32];
    __shared__ __align__(8) uint64_t mbar;
    DeviceCtx cx = make_ctx(red, &mbar, 0);
    yr::abs_mean_worker(cx, a, blockIdx.x);
}

__global__ void subbox_mean_kernel(yr::DecayTensor t, int dim, int axis, double* chunk) {
    __shared__ double red[32];
    __shared__ __align__(8) uint64_t mbar;
    DeviceCtx cx = make_ctx(red, &mbar, 0);
    yr::subbox_axis_mean_worker(cx, t, dim, axis, chunk, blockIdx.x);
}
__global__ void maxdiff_kernel(yr::DecayTensor t2, yr::DecayTensor t1, int dim, double* part) {
    __shared__ double red[32];
    __shared__ __align__(8) uint64_t mbar;
    DeviceCtx cx = make_ctx(red, &mbar, 0);
    yr::maxdiff_worker(cx, t2, t1, dim, part, blockIdx.x, gridDim.x);
}
__global__ void absmax_kernel(const double* v, long long n, double* part) {
    __shared__ double red[32];
    __shared__ __align__(8) uint64_t mbar;
    DeviceCtx cx = make_ctx(red, &mbar, 0);
    yr::absmax_worker(cx, v, n, part, blockIdx.x, gridDim.x);
}
__global__ void decay_decide_kernel(int mode, const double* chunk, long long n, const double* sup_part, long long nsup, const double* diff_part,
                                    long long ndiff, double prev_supnorm, double rel_tol, double abs_tol, yr::DecayOut* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) yr::decay_decide(mode, chunk, n, sup_part, nsup, diff_part, ndiff, prev_supnorm, rel_tol, abs_tol, out);
}

int cuda_fail(cudaError_t e, const char* what) {
    char buf[400];
    snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    return yr_api::fail(buf);
}

template <class K>
int set_smem(K kernel, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return e == cudaSuccess ? 0 : cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
}

}  // namespace

namespace yr_backend {

#define YR_STR2(x) #x
#define YR_STR(x) YR_STR2(x)
const char* build_info() {
    return "cuda sm_100a (nvcc " YR_STR(__CUDACC_VER_MAJOR__) "." YR_STR(__CUDACC_VER_MINOR__) "." YR_STR(__CUDACC_VER_BUILD__)
           ", host compiler " __VERSION__ ", -fmad=false)";
}

int device_limits(int32_t* sm_count, int64_t* max_smem_per_cta, int64_t* smem_per_sm) {
    int dev = 0, v = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if (sm_count) { cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); *sm_count = v; }
    if (max_smem_per_cta) { cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev); *max_smem_per_cta = v; }
    if (smem_per_sm) { cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev); *smem_per_sm = v; }
    return 0;
}


template <int D, class KC, class KW>
int launch_box_step(const yr::PhaseArgs<D>& pa, const yr_launch& l, void* stream, KC cta_kernel, KW warp_kernel, const char* what) {
    if (pa.n_list == 0) return 0;
    // never more workers than boxes of this step
    cudaError_t e;
    if (l.warp_per_box) {
        if (!l.ws_in_smem) return yr_api::fail("warp-per-box mode needs the workspace in shared memory");
        const int warps = l.threads / 32;
        const size_t stride = align16(sizeof(yr::BoxShared<D>)) + 16 + align16((size_t)l.ws_doubles * 8);
        const size_t smem = stride * (size_t)warps;
        if (smem > 48 * 1024 && set_smem(warp_kernel, smem)) return -1;
        unsigned long long want = (pa.n_list + warps - 1) / warps;
        const unsigned ctas = (unsigned)(want < (unsigned long long)l.ctas ? want : (unsigned long long)l.ctas);
        warp_kernel<<<ctas, l.threads, smem, (cudaStream_t)stream>>>(pa, (unsigned)stride);
    } else {
        const size_t smem = l.ws_in_smem ? (size_t)l.ws_doubles * 8 : 0;
        if (!l.ws_in_smem && !l.gws) return yr_api::fail("global workspace missing");
        if (smem > 48 * 1024 && set_smem(cta_kernel, smem)) return -1;
        const unsigned ctas = (unsigned)(pa.n_list < (unsigned long long)l.ctas ? pa.n_list : (unsigned long long)l.ctas);
        cta_kernel<<<ctas, l.threads, smem, (cudaStream_t)stream>>>(pa, l.ws_in_smem);
    }
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, what);
}

template <int D>
int launch_tensor_step(const yr::PhaseArgs<D>& pa, const yr_launch& l, void* stream) {
    return launch_box_step<D>(pa, l, stream, tensor_step_kernel<D>, tensor_step_kernel_warp<D>, "tensor_step launch");
}

template <int D>
int launch_subdivide_step(const yr::PhaseArgs<D>& pa, const yr_launch& l, void* stream) {
    return launch_box_step<D>(pa, l, stream, subdivide_step_kernel<D>, subdivide_step_kernel_warp<D>, "subdivide_step launch");
}

template <int D>
int launch_scalar_step(const yr::PhaseArgs<D>& pa, void* stream) {
    if (pa.n_list == 0) return 0;
    scalar_step_kernel<D><<<(unsigned)((pa.n_list + 127) / 128), 128, 0, (cudaStream_t)stream>>>(pa);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "scalar_step launch");
}

int read_words(const void* dev, unsigned long long* out, int count, void* stream) {
    static unsigned long long* pinned = nullptr;
    if (!pinned) {
        cudaError_t e0 = cudaHostAlloc(reinterpret_cast<void**>(&pinned), 256, cudaHostAllocDefault);
        if (e0 != cudaSuccess) return cuda_fail(e0, "cudaHostAlloc");
    }
    if (count > 32) return yr_api::fail("read_words: too many words");
    cudaError_t e = cudaMemcpyAsync(pinned, dev, 8 * (size_t)count, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "counter read-back");
    for (int i = 0; i < count; ++i) out[i] = pinned[i];
    return 0;
}

static long env_long(const char* name, long dflt) { const char* v = getenv(name); return v ? atol(v) : dflt; }

// Launch shape of one step: warp-per-box (4 independent warp-teams per CTA, up to 16 per SM) while a
// box fits a warp's slice of shared memory and there are enough boxes to fill the machine; CTA-per-box
// otherwise; the caller's global workspace when a box does not fit shared memory at all.
template <int D>
int plan_step(int subdivide, unsigned long long n_items, unsigned long long total, unsigned long long tensor,
              unsigned long long extent, int exact, const yr_launch& base, yr_launch* out) {
    static int sm = 0, smem_cta = 0, smem_sm = 0;
    if (!sm) {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&smem_cta, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    }
    static const long warp_min_items = env_long("YR_WARP_MIN_ITEMS", 1024), warps_per_sm = env_long("YR_WARP_TEAMS", 16);
    if (total < 1) total = 1;
    if (tensor < 1) tensor = 1;
    if (extent < 1) extent = 1;
    const unsigned long long ws = yr::workspace_doubles(D, total, tensor, extent, exact, subdivide);
    *out = base;
    out->ws_doubles = ws; out->gws = nullptr; out->gws_stride = 0;
    const size_t per_warp = align16(sizeof(yr::BoxShared<D>)) + 16 + align16((size_t)ws * 8);
    if ((long)n_items >= warp_min_items && per_warp + 1024 <= (size_t)smem_cta) {
        int warps = (int)(((size_t)smem_cta - 1024) / per_warp); if (warps > 4) warps = 4;
        const size_t per_cta = (size_t)warps * per_warp + 1024;
        long per_sm = (long)((size_t)smem_sm / per_cta); if (per_sm > warps_per_sm / warps) per_sm = warps_per_sm / warps; if (per_sm < 1) per_sm = 1;
        unsigned long long want = (n_items + warps - 1) / warps, cap = (unsigned long long)sm * per_sm;
        out->threads = 32 * warps; out->ctas = (int)(want < cap ? want : cap); out->ws_in_smem = 1; out->warp_per_box = 1;
        return 0;
    }
    out->warp_per_box = 0;
    out->threads = total >= 768 ? 128 : 64;
    const size_t need = (size_t)ws * 8 + sizeof(yr::BoxShared<D>) + 512;
    if (need <= (size_t)smem_cta) {
        long per_sm = (long)((size_t)smem_sm / (need + 1024)); if (per_sm > 2048 / out->threads) per_sm = 2048 / out->threads; if (per_sm > 32) per_sm = 32; if (per_sm < 1) per_sm = 1;
        unsigned long long cap = (unsigned long long)sm * per_sm;
        out->ctas = (int)(n_items < cap ? n_items : cap); out->ws_in_smem = 1;
        return 0;
    }
    if (!base.gws || base.gws_stride < ws) return yr_api::fail("box does not fit shared memory and no global workspace was provided");
    out->ws_in_smem = 0; out->gws = base.gws; out->gws_stride = base.gws_stride;
    out->ctas = (int)(n_items < (unsigned long long)base.ctas ? n_items : (unsigned long long)base.ctas);
    return 0;
}

int zero_bytes(void* dev, size_t bytes, void* stream) {
    cudaError_t e = cudaMemsetAsync(dev, 0, bytes, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : cuda_fail(e, "cudaMemsetAsync");
}

template <int D>
int launch_init(const yr::InitArgs<D>& a, const yr_launch& l, void* stream) {
    const size_t smem = l.ws_in_smem ? (size_t)l.ws_doubles * 8 : 0;
    if (!l.ws_in_smem && !l.gws) return yr_api::fail("global workspace missing");
    if (smem > 48 * 1024 && set_smem(init_kernel<D>, smem)) return -1;
    if (a.njobs == 0) return 0;
    init_kernel<D><<<l.ctas, l.threads, smem, (cudaStream_t)stream>>>(a, l.ws_in_smem);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "init_kernel launch");
}

template <int D>
int launch_bils(int64_t n, const double* lin, const double* c0, const double* sumabs, const double* errs,
                const int32_t* final_step, double* interval, int32_t* flags, void* stream) {
    if (n == 0) return 0;
    bils_kernel<D><<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(n, lin, c0, sumabs, errs, final_step, interval, flags);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "bils_kernel launch");
}

template <int D>
int launch_quad(int64_t n, const double* coeffs, const int64_t* off, const int32_t* shapes, const double* tol,
                int32_t nd_check, int32_t* out, void* stream) {
    if (n == 0) return 0;
    quad_kernel<D><<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        n, coeffs, reinterpret_cast<const long long*>(off), shapes, tol, nd_check, out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "quad_kernel launch");
}

static unsigned grid_for(long long total, int threads) {
    long long g = (total + threads - 1) / threads;
    const long long cap = 148ll * 8;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

int launch_dct1(const double* values, double* coeffs, int64_t outer, int64_t npts, int64_t inner, void* stream) {
    yr::Dct1Args a{values, coeffs, outer, npts, inner};
    const size_t smem = (size_t)(2 * (npts - 1)) * 8;
    if (smem > 200 * 1024) {
        // an axis beyond 12 800 points (the degree search doubles its guess up to 1e5, :280): cosine table in global memory
        double* table = nullptr;
        cudaError_t e0 = cudaMallocAsync(&table, smem, (cudaStream_t)stream);
        if (e0 != cudaSuccess) return cuda_fail(e0, "cudaMallocAsync (DCT table)");
        dct1_table_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(table, npts - 1);
        dct1_kernel_gtable<<<grid_for(outer * npts * inner, 256), 256, 0, (cudaStream_t)stream>>>(a, table);
        cudaError_t e1 = cudaGetLastError();
        cudaFreeAsync(table, (cudaStream_t)stream);
        return e1 == cudaSuccess ? 0 : cuda_fail(e1, "dct1_kernel_gtable launch");
    }
    if (smem > 48 * 1024 && set_smem(dct1_kernel, smem)) return -1;
    dct1_kernel<<<grid_for(outer * npts * inner, 256), 256, smem, (cudaStream_t)stream>>>(a);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "dct1_kernel launch");
}

int launch_poly_eval(const double* coef, const double* x, double* out, int64_t outer, int64_t ncoef, int64_t npts,
                     int64_t inner, int32_t basis, void* stream) {
    yr::PolyEvalArgs a{coef, x, out, outer, ncoef, npts, inner, basis};
    poly_eval_kernel<<<grid_for(outer * npts * inner, 256), 256, 0, (cudaStream_t)stream>>>(a);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "poly_eval_kernel launch");
}

int launch_upper_axis(const double* coef, const double* P, double* out, int64_t outer, int64_t n, int64_t inner, void* stream) {
    yr::UpperAxisArgs a{coef, P, out, outer, n, inner};
    upper_axis_kernel<<<grid_for(outer * n * inner, 256), 256, 0, (cudaStream_t)stream>>>(a);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "upper_axis_kernel launch");
}

int launch_abs_axis_mean(const double* t, double* out, double* supnorm, int64_t outer, int64_t n, int64_t inner, void* stream) {
    yr::AbsMeanArgs a{t, out, supnorm, outer, n, inner};
    abs_mean_kernel<<<(unsigned)n, 256, 0, (cudaStream_t)stream>>>(a);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "abs_mean_kernel launch");
}

// Degree-selection probe: scratch = [chunk: n][sup partials: 64][diff partials: 64] doubles (device), out on the device.
int launch_decay(int mode, int dim, int axis, const yr::DecayTensor& t, const yr::DecayTensor* prev, const double* values, int64_t nvalues,
                 double prev_supnorm, double rel_tol, double abs_tol, double* scratch, yr::DecayOut* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = t.keep[axis];
    double* chunk = scratch; double* sup_part = scratch + n; double* diff_part = sup_part + 64;
    subbox_mean_kernel<<<(unsigned)n, 128, 0, st>>>(t, dim, axis, chunk);
    absmax_kernel<<<64, 256, 0, st>>>(values, nvalues, sup_part);
    if (mode == 1) maxdiff_kernel<<<64, 256, 0, st>>>(t, *prev, dim, diff_part);
    decay_decide_kernel<<<1, 32, 0, st>>>(mode, chunk, n, sup_part, 64, diff_part, mode == 1 ? 64 : 0, prev_supnorm, rel_tol, abs_tol, out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "decay kernels launch");
}

}  // namespace yr_backend
