// CUDA backend (sm_100a) of the D-dimensional frontier pipeline (yr_fd.h), D = 3 and 4.
//
// Kernels:
//   fd_tensor_warp<D,OP,FMA>        4 independent 32-lane teams per CTA, each with its own slice of dynamic shared memory
//   fd_tensor_cta<D,OP,FMA,THREADS> one team of 64 / 256 threads per CTA (boxes that need more than a warp's slice; 128 on request)
//   fd_scalar_kernel<D>             thread per box
//   fd_import_kernel<D>, fd_sort_keys<D>, fd_pack_sorted<D>
// Grids are one resident wave; teams take boxes from a per-(op, class) cursor.  Compiled with -fmad=false.
#include <cuda_runtime.h>
#include <cuda_pipeline.h>
#include <cub/device/device_radix_sort.cuh>
#include <stdint.h>
#include <stdlib.h>
#include <vector>

// the per-box mathematics inline, so that its small arrays live in registers instead of a stack frame (as in yr_f2_kernels.cu)
#ifndef YR_FD_INLINE_MATH
#define YR_FD_INLINE_MATH 1
#endif
#if YR_FD_INLINE_MATH
#define YR_HDM static __host__ __device__ __forceinline__     // see yr_math.h
#endif
#include "yr_fd_host.h"
#include "yr_cuda_teams.cuh"

using namespace yr;
using namespace yr::fd;
using namespace yr_cuda;

namespace {

int cuda_fail(cudaError_t e, const char* what) {
    char buf[400];
    snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    yr_set_error(buf);
    return -1;
}

// copy the m = 0 subdivision matrices into shared memory, re-laid out for extent `snt` (<= table extent)
__device__ void stage_submats(const Tables& tab, double* smat, int snt) {
    const int nt = tab.nt;
    const int B = (snt + 3) >> 2;
    const int P = panel_doubles(snt);
    for (int h = 0; h < 2; ++h) {
        const double* src = tab.mat[0][h];
        double* dst = smat + h * P;
        for (int b = 0; b < B; ++b) {
            const int cnt = (snt - 4 * b) * 4;
            const double* s = src + panel_off(b, nt);
            double* d = dst + panel_off(b, snt);
            for (int e = threadIdx.x; e < cnt; e += blockDim.x) d[e] = s[e];
        }
    }
    __syncthreads();
}

template <int D, int OP, bool FMA>
__global__ void __launch_bounds__(128, 4) fd_tensor_warp(const Args<D> a, const unsigned int* __restrict__ list, unsigned long long n,
                                                         unsigned slice_doubles, int snt, unsigned long long* cursor) {
    extern __shared__ __align__(16) double dyn[];
    __shared__ uint64_t team_bars[4];
    __shared__ WalkState<D> walk[OP == kTSubdivide ? 4 : 1];
    double* smat = dyn;
    int smat_doubles = 0;
    if (OP == kTSubdivide && snt > 0) { smat_doubles = 2 * panel_doubles(snt); stage_submats(a.tab, smat, snt); }
    const int w = threadIdx.x >> 5;
    double* slice = dyn + smat_doubles + (size_t)w * slice_doubles;
    WarpTeam<32> tm = make_warp_team<32>(team_bars);
    unsigned long long macs = 0;
    auto grab = [&]() { unsigned long long v = 0; if (tm.tid == 0) v = atomicAdd(cursor, 1ull); return tm.bcast(v); };
    unsigned long long i = grab();
    while (i < n) {
        const unsigned int b = list[i];
        if (OP == kTRestrict) restrict_box<D, FMA>(tm, a, b, slice, macs);
        else subdivide_box<D, FMA>(tm, a, b, slice, snt > 0 ? smat : nullptr, snt > 0 ? smat + panel_doubles(snt) : nullptr, snt, macs,
                                   &walk[OP == kTSubdivide ? w : 0]);
        tm.sync();
        i = grab();
    }
    if (tm.tid == 0 && macs) atomicAdd(&a.ctl->restrict_macs, macs);
}

template <int D, int OP, bool FMA, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 64 ? 6 : (THREADS == 128 ? 3 : 1)) fd_tensor_cta(const Args<D> a, const unsigned int* __restrict__ list,
                                                                                unsigned long long n, int snt, unsigned long long* cursor) {
    extern __shared__ __align__(16) double dyn[];
    __shared__ CtaShared cta_sh;
    constexpr bool kPerWarp = (THREADS <= 128);
    __shared__ WalkState<D> walk[(OP == kTSubdivide && kPerWarp) ? THREADS / 32 : 1];
    double* smat = dyn;
    int smat_doubles = 0;
    if (OP == kTSubdivide && snt > 0) { smat_doubles = 2 * panel_doubles(snt); stage_submats(a.tab, smat, snt); }
    double* slice = dyn + smat_doubles;
    CtaTeam tm = make_cta_team(&cta_sh);
    unsigned long long macs = 0;
    for (;;) {
        unsigned long long i = 0;
        if (tm.tid == 0) i = atomicAdd(cursor, 1ull);
        i = tm.bcast(i);
        if (i >= n) break;
        const unsigned int b = list[i];
        if (OP == kTRestrict) restrict_box<D, FMA>(tm, a, b, slice, macs);
        else subdivide_box<D, FMA>(tm, a, b, slice, snt > 0 ? smat : nullptr, snt > 0 ? smat + panel_doubles(snt) : nullptr, snt, macs,
                                   &walk[(OP == kTSubdivide && kPerWarp) ? (threadIdx.x >> 5) : 0], kPerWarp);
        __syncthreads();
    }
    if (tm.tid == 0 && macs) atomicAdd(&a.ctl->restrict_macs, macs);
}

// queue a warp's boxes with one atomic per (op, class) present in the warp, and flush its statistics
template <int D>
__device__ void commit_warp(const Args<D>& a, unsigned int b, const Decision& dc) {
    const int lane = threadIdx.x & 31;
    const int key = dc.op == kTNone ? -1 : op_slot(dc.op) * kClasses + dc.cls;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (key >= 0) {
        const int leader = __ffs(peers) - 1;
        const int rank = __popc(peers & ((1u << lane) - 1u));
        const unsigned need = __reduce_max_sync(peers, (unsigned)dc.need);
        const unsigned ext = __reduce_max_sync(peers, (unsigned)dc.ext);
        unsigned long long base = 0;
        const int slot = op_slot(dc.op);
        if (lane == leader) {
            base = atomicAdd(&a.ctl->n_list[slot][dc.cls], (unsigned long long)__popc(peers));
            atomicMax(&a.ctl->need_max[slot][dc.cls], (unsigned long long)need);
            if (slot == 1) atomicMax(&a.ctl->ext_max[dc.cls], (unsigned long long)ext);
        }
        base = __shfl_sync(peers, base, leader);
        a.lists[slot][dc.cls][base + rank] = b;
    }
    const unsigned long long bsum = warp_add(dc.op == kTNone ? 0ull : dc.bound);
    unsigned long long v[7] = {dc.boxes, dc.zooms, dc.dconst, dc.dquad, dc.dzoom, dc.coeffs, dc.subdiv};
#pragma unroll
    for (int q = 0; q < 7; ++q) v[q] = warp_add(v[q]);
    const unsigned long long mx = warp_max(dc.maxlevel);
    if (lane == 0) {
        Ctl* c = a.ctl;
        if (bsum) atomicAdd(&c->out_bound, bsum);
        if (v[0]) atomicAdd(&c->boxes, v[0]);
        if (v[1]) atomicAdd(&c->zooms, v[1]);
        if (v[2]) atomicAdd(&c->dconst, v[2]);
        if (v[3]) atomicAdd(&c->dquad, v[3]);
        if (v[4]) atomicAdd(&c->dzoom, v[4]);
        if (v[5]) atomicAdd(&c->entry_coeffs, v[5]);
        if (v[6]) atomicAdd(&c->subdivisions, v[6]);
        if (mx) atomicMax(&c->max_level, mx);
    }
}

template <int D>
__global__ void __launch_bounds__(128) fd_scalar_kernel(const Args<D> a, const yr_fd_backend::Segments sg) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    Decision dc;
    memset(&dc, 0, sizeof(dc));
    bool have = false;
    unsigned int b = 0;
#pragma unroll
    for (int c = 0; c < kClasses; ++c) {
        if (!have) { if (i < sg.n[c]) { b = sg.list[c][i]; have = true; } else i -= sg.n[c]; }
    }
    if (!have) {
        const unsigned long long nchild = a.ctl->n_boxes - sg.child_lo;
        if (i < nchild) { b = (unsigned int)(sg.child_lo + i); have = true; }
    }
    if (have) scalar_step<D>(a, b, dc, DevAt());
    commit_warp<D>(a, b, dc);
}

template <int D>
__global__ void __launch_bounds__(128) fd_import_kernel(const Args<D> a, unsigned long long n) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    Decision dc;
    memset(&dc, 0, sizeof(dc));
    if (i < n) import_box<D>(a, (unsigned int)i, dc, DevAt());
    commit_warp<D>(a, (unsigned int)i, dc);
}

struct Dev { int sm = 0, smem_cta = 0; };
const Dev& dev() {
    static Dev d;
    if (!d.sm) {
        int id = 0; cudaGetDevice(&id);
        cudaDeviceGetAttribute(&d.sm, cudaDevAttrMultiProcessorCount, id);
        cudaDeviceGetAttribute(&d.smem_cta, cudaDevAttrMaxSharedMemoryPerBlockOptin, id);
    }
    return d;
}
constexpr size_t kStaticReserve = 2048;          // team barriers, reduction scratch, the walk states of 4 teams (yr_fd.h: WalkState)
template <class K>
int ensure_smem(K kernel, size_t bytes) {
    static std::vector<const void*> done;
    if (bytes + kStaticReserve <= 48 * 1024) return 0;          // (static shared memory counts against the default 48 KB limit)
    const void* key = reinterpret_cast<const void*>(kernel);
    for (const void* k : done) if (k == key) return 0;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dev().smem_cta - (int)kStaticReserve);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
    done.push_back(key);
    return 0;
}
template <class K>
int resident_ctas(K kernel, int threads, size_t smem) {
    int r = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, kernel, threads, smem) != cudaSuccess || r < 1) { cudaGetLastError(); r = 1; }
    return r;
}

template <int D, int OP, bool FMA>
int launch_tensor_t(const Args<D>& a, int cls, const unsigned int* list, unsigned long long n, unsigned long long need, int ext, cudaStream_t st) {
    const Dev& d = dev();
    unsigned long long* cursor = &a.ctl->cursor[op_slot(OP)][cls];
    const int snt = (OP == kTSubdivide) ? (ext < 4 ? 4 : ext) : 0;
    const size_t smat = (OP == kTSubdivide) ? (size_t)2 * panel_doubles(snt) * 8 : 0;
    const size_t slice = ((size_t)need + 1) & ~(size_t)1;
    if (cls < kWarpClasses) {
        // as many 32-lane teams per CTA as fit (4 at most)
        int teams = 4;
        while (teams > 1 && smat + teams * slice * 8 > (size_t)d.smem_cta - kStaticReserve) teams >>= 1;
        const size_t smem = smat + teams * slice * 8;
        if (smem > (size_t)d.smem_cta - kStaticReserve) { yr_set_error("fd: warp-team launch exceeds shared memory"); return -1; }
        const unsigned long long want = (n + teams - 1) / teams;
        if (ensure_smem(fd_tensor_warp<D, OP, FMA>, smem)) return -1;
        const unsigned long long cap = (unsigned long long)d.sm * resident_ctas(fd_tensor_warp<D, OP, FMA>, 32 * teams, smem);
        fd_tensor_warp<D, OP, FMA><<<(unsigned)(want < cap ? want : cap), 32 * teams, smem, st>>>(a, list, n, (unsigned)slice, snt, cursor);
    } else {
        const int threads = class_threads(cls);
        const size_t smem = smat + slice * 8;
        if (smem > (size_t)d.smem_cta - kStaticReserve) { yr_set_error("fd: CTA-team launch exceeds shared memory"); return -1; }
        // boxes beyond a warp's slice run on TWO warps (64 threads), not four: every phase outside the sweeps is serial per box
        // and each pass costs a CTA barrier (measured: 3-D degree 10 139.6 -> 138.0 ms, 4-D degree 4 116.8 -> 110.2 ms together
        // with the lower warp-team limit for SUBDIVIDE, profiles/r02_fd.txt).  YR_FD_CTA128=1 restores four warps.
        static const bool cta64 = getenv("YR_FD_CTA128") == nullptr;
        if (threads == 128 && cta64) {
            if (ensure_smem(fd_tensor_cta<D, OP, FMA, 64>, smem)) return -1;
            const unsigned long long cap = (unsigned long long)d.sm * resident_ctas(fd_tensor_cta<D, OP, FMA, 64>, 64, smem);
            fd_tensor_cta<D, OP, FMA, 64><<<(unsigned)(n < cap ? n : cap), 64, smem, st>>>(a, list, n, snt, cursor);
        } else if (threads == 128) {
            if (ensure_smem(fd_tensor_cta<D, OP, FMA, 128>, smem)) return -1;
            const unsigned long long cap = (unsigned long long)d.sm * resident_ctas(fd_tensor_cta<D, OP, FMA, 128>, 128, smem);
            fd_tensor_cta<D, OP, FMA, 128><<<(unsigned)(n < cap ? n : cap), 128, smem, st>>>(a, list, n, snt, cursor);
        } else {
            if (ensure_smem(fd_tensor_cta<D, OP, FMA, 256>, smem)) return -1;
            const unsigned long long cap = (unsigned long long)d.sm * resident_ctas(fd_tensor_cta<D, OP, FMA, 256>, 256, smem);
            fd_tensor_cta<D, OP, FMA, 256><<<(unsigned)(n < cap ? n : cap), 256, smem, st>>>(a, list, n, snt, cursor);
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        char what[200];
        snprintf(what, sizeof(what), "fd tensor launch (dim %d op %d class %d, %llu boxes, need %llu doubles, extent %d)", D, OP, cls, n, need, ext);
        return cuda_fail(e, what);
    }
    return 0;
}

// ---- reporting order: stable LSD radix sort of the record indices by path_lo, path_hi, system (CUB) ----
template <int D>
__global__ void fd_sort_keys(const ResultRec<D>* res, const unsigned int* idx, unsigned long long n, int which,
                             unsigned long long* k64, unsigned int* k32, unsigned int* iota) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int r = idx ? idx[i] : (unsigned int)i;
    if (iota) iota[i] = (unsigned int)i;
    if (which == 0) k64[i] = res[r].path_lo;
    else if (which == 1) k64[i] = res[r].path_hi;
    else k32[i] = (unsigned int)res[r].system;
}

template <int D> struct SortedRec { uint32_t index; int32_t system; uint32_t flags; int32_t collector; double root[D]; };

template <int D>
__global__ void fd_pack_sorted(const ResultRec<D>* res, const unsigned int* order, unsigned long long n, SortedRec<D>* out) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int r = order[i];
    SortedRec<D> o;
    o.index = r; o.system = res[r].system; o.flags = res[r].flags; o.collector = res[r].collector;
    for (int d = 0; d < D; ++d) o.root[d] = res[r].root[d];
    out[i] = o;
}

}  // namespace

namespace yr_fd_backend {
namespace be = yr_f2_backend;

template <int D>
int launch_import(const Args<D>& a, unsigned long long n, void* stream) {
    fd_import_kernel<D><<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a, n);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "fd_import_kernel launch");
}

template <int D>
int launch_tensor(const Args<D>& a, int op, int cls, const unsigned int* list, unsigned long long n, unsigned long long need, int ext, void* stream) {
    cudaStream_t st = (cudaStream_t)be::fan_stream(stream);
    const bool fma = a.prm.use_fma != 0;
    if (op == kTRestrict) return fma ? launch_tensor_t<D, kTRestrict, true>(a, cls, list, n, need, ext, st) : launch_tensor_t<D, kTRestrict, false>(a, cls, list, n, need, ext, st);
    return fma ? launch_tensor_t<D, kTSubdivide, true>(a, cls, list, n, need, ext, st) : launch_tensor_t<D, kTSubdivide, false>(a, cls, list, n, need, ext, st);
}

template <int D>
int launch_scalar(const Args<D>& a, const Segments& s, void* stream) {
    unsigned long long total = s.child_bound;
    for (int c = 0; c < kClasses; ++c) total += s.n[c];
    if (!total) return 0;
    fd_scalar_kernel<D><<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a, s);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "fd_scalar_kernel launch");
}

template <int D>
void* sort_results(const void* results, unsigned long long n, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const ResultRec<D>* res = static_cast<const ResultRec<D>*>(results);
    const size_t n_ = (size_t)n;
    unsigned long long* k64a = (unsigned long long*)be::dev_alloc(n_ * 8, stream); unsigned long long* k64b = (unsigned long long*)be::dev_alloc(n_ * 8, stream);
    unsigned int* k32a = (unsigned int*)be::dev_alloc(n_ * 4, stream); unsigned int* k32b = (unsigned int*)be::dev_alloc(n_ * 4, stream);
    unsigned int* ia = (unsigned int*)be::dev_alloc(n_ * 4, stream); unsigned int* ib = (unsigned int*)be::dev_alloc(n_ * 4, stream);
    void* result = nullptr;
    if (k64a && k64b && k32a && k32b && ia && ib) {
        const unsigned grid = (unsigned)((n + 255) / 256);
        size_t tb64 = 0, tb32 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb64, k64a, k64b, ia, ib, (int)n, 0, 64, st);
        cub::DeviceRadixSort::SortPairs(nullptr, tb32, k32a, k32b, ia, ib, (int)n, 0, 32, st);
        void* tmp = be::dev_alloc(tb64 > tb32 ? tb64 : tb32, stream);
        if (tmp) {
            fd_sort_keys<D><<<grid, 256, 0, st>>>(res, nullptr, n, 0, k64a, nullptr, ia);
            cub::DeviceRadixSort::SortPairs(tmp, tb64, k64a, k64b, ia, ib, (int)n, 0, 64, st);
            fd_sort_keys<D><<<grid, 256, 0, st>>>(res, ib, n, 1, k64a, nullptr, nullptr);
            cub::DeviceRadixSort::SortPairs(tmp, tb64, k64a, k64b, ib, ia, (int)n, 0, 64, st);
            fd_sort_keys<D><<<grid, 256, 0, st>>>(res, ia, n, 2, nullptr, k32a, nullptr);
            cub::DeviceRadixSort::SortPairs(tmp, tb32, k32a, k32b, ia, ib, (int)n, 0, 32, st);
            be::dev_free(tmp, stream);
            SortedRec<D>* packed = (SortedRec<D>*)be::dev_alloc(n_ * sizeof(SortedRec<D>), stream);
            if (packed) {
                fd_pack_sorted<D><<<grid, 256, 0, st>>>(res, ib, n, packed);
                if (cudaGetLastError() == cudaSuccess) result = packed; else be::dev_free(packed, stream);
            }
        }
    }
    be::dev_free(k64a, stream); be::dev_free(k64b, stream); be::dev_free(k32a, stream); be::dev_free(k32b, stream); be::dev_free(ia, stream); be::dev_free(ib, stream);
    return result;
}

template int launch_import<3>(const Args<3>&, unsigned long long, void*);
template int launch_import<4>(const Args<4>&, unsigned long long, void*);
template int launch_tensor<3>(const Args<3>&, int, int, const unsigned int*, unsigned long long, unsigned long long, int, void*);
template int launch_tensor<4>(const Args<4>&, int, int, const unsigned int*, unsigned long long, unsigned long long, int, void*);
template int launch_scalar<3>(const Args<3>&, const Segments&, void*);
template int launch_scalar<4>(const Args<4>&, const Segments&, void*);
template void* sort_results<3>(const void*, unsigned long long, void*);
template void* sort_results<4>(const void*, unsigned long long, void*);

}  // namespace yr_fd_backend

int yr_fd_solve(const yr_frontier_desc* d, yr_frontier_out* out, void* stream) {
    if (d->dim == 3) return yr_fd_host::solve<3>(d, out, stream);
    return yr_fd_host::solve<4>(d, out, stream);
}
int yr_fd_fits(int dim, unsigned long long max_extent, unsigned long long max_tensor) {
    if (dim == 3) return yr_fd_host::fits<3>(max_extent, max_tensor);
    if (dim == 4) return yr_fd_host::fits<4>(max_extent, max_tensor);
    return 0;
}
int yr_fd_trim(void* stream) { yr_fd_host::trim_workspace<3>(stream); yr_fd_host::trim_workspace<4>(stream); return 0; }
