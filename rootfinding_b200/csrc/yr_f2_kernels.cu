// CUDA backend (sm_100a) of the 2-D frontier pipeline: yr_frontier_solve / yr_frontier_release / yr_fetch.
//
// Kernels (see yr_f2.h for what a round is):
//   f2_tensor_warp<OP,FMA>   4 independent warp teams per CTA, each with its own slice of dynamic shared memory
//   f2_tensor_cta<OP,FMA>    one team of 128 / 256 threads per CTA
//   f2_scalar_kernel         thread per box
//   f2_import_kernel, f2_tables_kernel
// Grids are sized to the machine: teams = min(boxes, SMs x resident teams), every team strides the list.
// Compiled with -fmad=false like the rest of the library.
#include <cuda_runtime.h>
#include <cuda_pipeline.h>
#include <cub/device/device_radix_sort.cuh>
#include <stdint.h>
#include <stdlib.h>
#include <time.h>

#ifndef YR_F2_INLINE_MATH
#define YR_F2_INLINE_MATH 1
#endif
#if YR_F2_INLINE_MATH
#define YR_HDM static __host__ __device__ __forceinline__     // see yr_math.h
#endif
#include "yr_f2_host.h"
#include "yr_cuda_teams.cuh"

using namespace yr;
using namespace yr::f2;
using namespace yr_cuda;

int yr_fd_solve(const yr_frontier_desc* d, yr_frontier_out* out, void* stream);
int yr_fd_fits(int dim, unsigned long long max_extent, unsigned long long max_tensor);
int yr_fd_trim(void* stream);

namespace {

// copy the m = 0 subdivision matrices into shared memory, re-laid out for extent `snt` (<= table extent)
__device__ void stage_submats(const Args& a, double* smat, int snt) {
    const int nt = a.tab.nt;
    const int B = (snt + 3) >> 2;
    const int P = panel_doubles(snt);
    for (int h = 0; h < 2; ++h) {
        const double* src = a.tab.mat[0][h];
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

template <int OP, bool FMA, int LANES>
#ifndef YR_F2_WARP_MINB
#define YR_F2_WARP_MINB 5
#endif
__global__ void __launch_bounds__(128, YR_F2_WARP_MINB) f2_tensor_warp(const Args a, const unsigned int* __restrict__ list, unsigned long long n,
                                                      unsigned slice_doubles, int snt, unsigned long long* cursor) {
    extern __shared__ __align__(16) double dyn[];
    constexpr int kTeams = 128 / LANES;
    double* smat = dyn;
    int smat_doubles = 0;
    if (OP == kTSubdivide && snt > 0) { smat_doubles = 2 * panel_doubles(snt); stage_submats(a, smat, snt); }
    __shared__ uint64_t team_bars[8];
    const int w = threadIdx.x / LANES;
    double* slice = dyn + smat_doubles + (size_t)w * slice_doubles;
    WarpTeam<LANES> tm = make_warp_team<LANES>(team_bars);
    // The grid is one resident wave; teams take boxes from the queue one at a time (boxes of a class still differ in
    // cost), always holding the NEXT position already so that its records can be prefetched.
    unsigned long long macs = 0;
    auto grab = [&]() { unsigned long long v = 0; if (tm.tid == 0) v = atomicAdd(cursor, 1ull); return tm.bcast(v); };
    unsigned long long i = grab();
    while (i < n) {
        const unsigned long long nxt = grab();
        const unsigned int b = list[i];
        if (nxt < n && tm.tid < 7) {
            // the next box's records (BoxRec 3 lines, Work 3-4 lines) start their trip from HBM now
            const unsigned int nb = list[nxt];
            const char* p = tm.tid < 3 ? reinterpret_cast<const char*>(a.recs + nb) + 128 * tm.tid
                                       : reinterpret_cast<const char*>(a.work + nb) + 128 * (tm.tid - 3);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        }
        if (OP == kTRestrict) restrict_box_w<FMA>(tm, a, b, slice, macs);
        else subdivide_box_w<FMA>(tm, a, b, slice, snt > 0 ? smat : nullptr, snt > 0 ? smat + panel_doubles(snt) : nullptr, snt, macs);
        tm.sync();
        i = nxt;
    }
    if (tm.tid == 0 && macs) atomicAdd(&a.ctl->restrict_macs, macs);
}

#ifndef YR_F2_CTA128_MINB
#define YR_F2_CTA128_MINB 5
#endif
template <int OP, bool FMA, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 128 ? YR_F2_CTA128_MINB : 2) f2_tensor_cta(const Args a, const unsigned int* __restrict__ list, unsigned long long n, int snt,
                                                     unsigned long long* cursor) {
    extern __shared__ __align__(16) double dyn[];
    __shared__ CtaShared cta_sh;
    double* smat = dyn;
    int smat_doubles = 0;
    if (OP == kTSubdivide && snt > 0) { smat_doubles = 2 * panel_doubles(snt); stage_submats(a, smat, snt); }
    double* slice = dyn + smat_doubles;
    CtaTeam tm = make_cta_team(&cta_sh);
    unsigned long long macs = 0;
    for (;;) {
        unsigned long long i = 0;
        if (tm.tid == 0) i = atomicAdd(cursor, 1ull);
        i = tm.bcast(i);
        if (i >= n) break;
        const unsigned int b = list[i];
        if (OP == kTRestrict) restrict_box<FMA>(tm, a, b, slice, macs);
        else subdivide_box<FMA>(tm, a, b, slice, snt > 0 ? smat : nullptr, snt > 0 ? smat + panel_doubles(snt) : nullptr, snt, macs);
        __syncthreads();
    }
    if (tm.tid == 0 && macs) atomicAdd(&a.ctl->restrict_macs, macs);
}

// queue a warp's boxes with one atomic per (op, class) present in the warp, and flush its statistics
__device__ void commit_warp(const Args& a, unsigned int b, const Decision& dc) {
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

#ifndef YR_F2_SCALAR_MINB
#define YR_F2_SCALAR_MINB 1
#endif
__global__ void __launch_bounds__(128, YR_F2_SCALAR_MINB) f2_scalar_kernel(const Args a, const yr_f2_backend::Segments sg) {
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
    if (have) {
        // the step touches most of both records, one dependent access after another: pull all six lines towards
        // the SM first so the chain pays one memory round trip instead of many
        const char* pr = reinterpret_cast<const char*>(a.recs + b);
        const char* pw = reinterpret_cast<const char*>(a.work + b);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pr + 128 * q));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pw + 128 * q));
        }
        scalar_step(a, b, dc, DevAt());
    }
    commit_warp(a, b, dc);
}

__global__ void __launch_bounds__(128) f2_import_kernel(const Args a, unsigned long long n) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    Decision dc;
    memset(&dc, 0, sizeof(dc));
    if (i < n) import_box(a, (unsigned int)i, dc, DevAt());
    commit_warp(a, (unsigned int)i, dc);
}

// ---- records of selected systems only (yr_gather_by_system) ----
__global__ void gather_by_system_kernel(const unsigned char* table, unsigned long long n, unsigned rec_bytes, unsigned system_off,
                                        const unsigned char* sys_flags, unsigned long long nsys, unsigned int* out_index,
                                        unsigned char* out_recs, unsigned long long cap, unsigned long long* out_n) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned char* rec = table + i * rec_bytes;
    const int sys = *reinterpret_cast<const int*>(rec + system_off);
    if (sys < 0 || (unsigned long long)sys >= nsys || !sys_flags[sys]) return;
    const unsigned long long pos = atomicAdd(out_n, 1ull);
    if (pos >= cap) return;
    out_index[pos] = (unsigned int)i;
    const unsigned long long* s8 = reinterpret_cast<const unsigned long long*>(rec);
    unsigned long long* d8 = reinterpret_cast<unsigned long long*>(out_recs + pos * rec_bytes);
    for (unsigned q = 0; q < rec_bytes / 8; ++q) d8[q] = s8[q];
}

// ---- records of several ranks in one table (yr_frontier_adopt) ----
__global__ void remap_links_kernel(ResultRec<2>* res, unsigned long long nr, NodeRec<2>* nodes, unsigned long long nn, const yr_f2_backend::RankBases rb) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int mask = (1 << rb.shift) - 1;
    if (i < nr) {
        const int p = res[i].parent_node, c = res[i].collector;
        if (p >= 0) res[i].parent_node = (int)(rb.node[p >> rb.shift] + (unsigned long long)(p & mask));
        if (c >= 0) res[i].collector = (int)(rb.res[c >> rb.shift] + (unsigned long long)(c & mask));
    }
    if (i < nn) {
        const int p = nodes[i].parent_node;
        if (p >= 0) nodes[i].parent_node = (int)(rb.node[p >> rb.shift] + (unsigned long long)(p & mask));
    }
}

// ---- moving queued boxes between frontiers (yr_frontier_export / _import) ----
__global__ void __launch_bounds__(128) f2_export_kernel(const Args a, const yr_f2_host::XTake take, unsigned long long n, unsigned char* xbuf,
                                                        const double* __restrict__ data_in) {
    const unsigned long long j = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;     // warp per box
    const int lane = threadIdx.x & 31;
    if (j >= n) return;
    const unsigned int b = yr_f2_host::export_box_of(a, take, j);
    int sz0;
    const int tot = yr_f2_host::export_doubles(a, b, sz0);
    unsigned long long at = 0;
    if (lane == 0) at = atomicAdd(&reinterpret_cast<yr_f2_host::XHeader*>(xbuf)->data_doubles, (unsigned long long)tot);
    at = __shfl_sync(0xffffffffu, at, 0);
    yr_f2_host::export_one(a, b, j, n, xbuf, data_in, lane, 32, at);
    __syncwarp();
    if (lane == 0) yr_f2_host::export_fix_offsets(j, n, xbuf, at, sz0);
}

__global__ void __launch_bounds__(128) f2_import_x_kernel(const Args a, unsigned long long n, const unsigned char* __restrict__ xbuf,
                                                          unsigned long long slot_base, unsigned long long data_base) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    Decision dc;
    memset(&dc, 0, sizeof(dc));
    if (i < n) yr_f2_host::import_one(a, i, n, xbuf, (unsigned int)(slot_base + i), data_base, dc, DevAt());
    commit_warp(a, (unsigned int)(slot_base + i), dc);
}

// the four level-independent matrices: CTA m builds (mid, half) = (m >> 1, m & 1)
__global__ void __launch_bounds__(256) f2_tables_kernel(double* mats, int* rows, int nt) {
    __shared__ CtaShared cta_sh;
    CtaTeam tm = make_cta_team(&cta_sh);
    const int m = blockIdx.x;
    const double mid = (m >> 1) ? kFirstSplit : 0.0;
    const double alpha = (mid + 1) / 2, beta = (mid - 1) / 2;           // :1089
    const bool upper = m & 1;
    build_panels(tm, mats + (size_t)m * panel_doubles(nt), rows + (size_t)m * nt, nt, upper ? -beta : alpha, upper ? alpha : beta);
}

int cuda_fail(cudaError_t e, const char* what) {
    char buf[400];
    snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    yr_set_error(buf);
    return -1;
}

struct Dev { int sm = 0, smem_cta = 0, smem_sm = 0; };
const Dev& dev() {
    static Dev d;
    if (!d.sm) {
        int id = 0; cudaGetDevice(&id);
        cudaDeviceGetAttribute(&d.sm, cudaDevAttrMultiProcessorCount, id);
        cudaDeviceGetAttribute(&d.smem_cta, cudaDevAttrMaxSharedMemoryPerBlockOptin, id);
        cudaDeviceGetAttribute(&d.smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, id);
        // keep freed frontier buffers in the stream-ordered pool across the per-round synchronisations
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, id) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    return d;
}

// opt in to large dynamic shared memory, once per kernel (static shared memory counts against the CTA limit)
constexpr size_t kStaticReserve = 1024;
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

// CTAs of `kernel` one SM really holds at this shared-memory size (registers, shared memory, warps): the grid is
// exactly one resident wave
template <class K>
int resident_ctas(K kernel, int threads, size_t smem) {
    int r = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, kernel, threads, smem) != cudaSuccess || r < 1) { cudaGetLastError(); r = 1; }
    return r;
}

template <int OP, bool FMA>
int launch_tensor_t(const Args& a, int cls, const unsigned int* list, unsigned long long n, unsigned long long need, int ext, cudaStream_t st) {
    const Dev& d = dev();
    unsigned long long* cursor = &a.ctl->cursor[op_slot(OP)][cls];
    const int snt = (OP == kTSubdivide) ? (ext < 4 ? 4 : ext) : 0;
    const size_t smat = (OP == kTSubdivide) ? (size_t)2 * panel_doubles(snt) * 8 : 0;
    const size_t slice = ((size_t)need + 1) & ~(size_t)1;
    if (cls < kWarpClasses) {
        const int teams = cls < kLane16Classes ? 8 : 4;          // two 16-lane teams per warp for the small classes
        const size_t smem = smat + teams * slice * 8;
        if (smem > (size_t)d.smem_cta - kStaticReserve) { yr_set_error("f2: warp-team launch exceeds shared memory"); return -1; }
        const unsigned long long want = (n + teams - 1) / teams;
        if (cls < kLane16Classes) {
            if (ensure_smem(f2_tensor_warp<OP, FMA, 16>, smem)) return -1;
            const unsigned long long cap = (unsigned long long)d.sm * resident_ctas(f2_tensor_warp<OP, FMA, 16>, 128, smem);
            f2_tensor_warp<OP, FMA, 16><<<(unsigned)(want < cap ? want : cap), 128, smem, st>>>(a, list, n, (unsigned)slice, snt, cursor);
        } else {
            if (ensure_smem(f2_tensor_warp<OP, FMA, 32>, smem)) return -1;
            const unsigned long long cap = (unsigned long long)d.sm * resident_ctas(f2_tensor_warp<OP, FMA, 32>, 128, smem);
            f2_tensor_warp<OP, FMA, 32><<<(unsigned)(want < cap ? want : cap), 128, smem, st>>>(a, list, n, (unsigned)slice, snt, cursor);
        }
    } else {
        const int threads = class_threads(cls);
        const size_t smem = smat + slice * 8;
        if (smem > (size_t)d.smem_cta - kStaticReserve) { yr_set_error("f2: CTA-team launch exceeds shared memory"); return -1; }
        if (threads == 128) {
            if (ensure_smem(f2_tensor_cta<OP, FMA, 128>, smem)) return -1;
            const unsigned long long cap = (unsigned long long)d.sm * resident_ctas(f2_tensor_cta<OP, FMA, 128>, 128, smem);
            f2_tensor_cta<OP, FMA, 128><<<(unsigned)(n < cap ? n : cap), 128, smem, st>>>(a, list, n, snt, cursor);
        } else {
            if (ensure_smem(f2_tensor_cta<OP, FMA, 256>, smem)) return -1;
            const unsigned long long cap = (unsigned long long)d.sm * resident_ctas(f2_tensor_cta<OP, FMA, 256>, 256, smem);
            f2_tensor_cta<OP, FMA, 256><<<(unsigned)(n < cap ? n : cap), 256, smem, st>>>(a, list, n, snt, cursor);
        }
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "f2 tensor launch");
}

struct Timer {
    int enabled = 0;
    std::vector<cudaEvent_t> ev[2];          // begin/end pairs per family
    size_t used[2] = {0, 0};
} g_timer;

}  // namespace

namespace yr_f2_backend {

void* dev_alloc(size_t bytes, void* stream) {
    void* p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 16, (cudaStream_t)stream);
    if (e != cudaSuccess) { cuda_fail(e, "cudaMallocAsync"); return nullptr; }
    return p;
}
void dev_free(void* p, void* stream) { if (p) cudaFreeAsync(p, (cudaStream_t)stream); }
int dev_copy(void* dst, const void* src, size_t bytes, void* stream) {
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : cuda_fail(e, "cudaMemcpyAsync (d2d)");
}
int dev_zero(void* p, size_t bytes, void* stream) {
    cudaError_t e = cudaMemsetAsync(p, 0, bytes, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : cuda_fail(e, "cudaMemsetAsync");
}
int to_host(void* dst, const void* src, size_t bytes, void* stream) {
    static void* pinned = nullptr; static size_t cap = 0;
    if (bytes <= 4096) {
        if (!pinned) { cudaError_t e0 = cudaHostAlloc(&pinned, 4096, cudaHostAllocDefault); if (e0 != cudaSuccess) return cuda_fail(e0, "cudaHostAlloc"); cap = 4096; }
        cudaError_t e = cudaMemcpyAsync(pinned, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
        if (e != cudaSuccess) return cuda_fail(e, "control block read-back");
        memcpy(dst, pinned, bytes);
        return 0;
    }
    (void)cap;
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    return e == cudaSuccess ? 0 : cuda_fail(e, "device -> host copy");
}

int to_device(void* dst, const void* src, size_t bytes, void* stream) {
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    return e == cudaSuccess ? 0 : cuda_fail(e, "host -> device copy");
}

int launch_export(const Args& a, const yr_f2_host::XTake& take, unsigned long long n, unsigned char* xbuf, const double* data_in, void* stream) {
    f2_export_kernel<<<(unsigned)((n * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a, take, n, xbuf, data_in);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "f2_export_kernel launch");
}

int launch_import_x(const Args& a, unsigned long long n, const unsigned char* xbuf, unsigned long long slot_base, unsigned long long data_base, void* stream) {
    f2_import_x_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a, n, xbuf, slot_base, data_base);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "f2_import_x_kernel launch");
}

int launch_tables(double* mats, int* rows, int nt, void* stream) {
    f2_tables_kernel<<<4, 256, 0, (cudaStream_t)stream>>>(mats, rows, nt);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "f2_tables_kernel launch");
}

int launch_import(const Args& a, unsigned long long n, void* stream) {
    f2_import_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a, n);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "f2_import_kernel launch");
}

// ---- a round's tensor launches fan out over a few streams and join before the scalar step ----
namespace {
constexpr int kFanStreams = 6;
struct Fan { cudaStream_t st[kFanStreams]; cudaEvent_t done[kFanStreams]; cudaEvent_t start; bool ready = false, used[kFanStreams]; int next = 0; bool active = false; } g_fan;
}
void round_fork(void* stream) {
    static const bool enabled = getenv("YR_F2_ONE_STREAM") == nullptr && getenv("YR_F2_TRACE") == nullptr;
    g_fan.active = false;
    if (!enabled) return;
    if (!g_fan.ready) {
        for (int i = 0; i < kFanStreams; ++i) { cudaStreamCreateWithFlags(&g_fan.st[i], cudaStreamNonBlocking); cudaEventCreateWithFlags(&g_fan.done[i], cudaEventDisableTiming); }
        cudaEventCreateWithFlags(&g_fan.start, cudaEventDisableTiming);
        g_fan.ready = true;
    }
    cudaEventRecord(g_fan.start, (cudaStream_t)stream);
    for (int i = 0; i < kFanStreams; ++i) g_fan.used[i] = false;
    g_fan.next = 0; g_fan.active = true;
}
void round_join(void* stream) {
    if (!g_fan.active) return;
    for (int i = 0; i < kFanStreams; ++i)
        if (g_fan.used[i]) { cudaEventRecord(g_fan.done[i], g_fan.st[i]); cudaStreamWaitEvent((cudaStream_t)stream, g_fan.done[i], 0); }
    g_fan.active = false;
}

// the stream the next tensor launch of the current round goes to (between round_fork and round_join: one of the fan streams)
void* fan_stream(void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (g_fan.active) {
        const int i = g_fan.next; g_fan.next = (g_fan.next + 1) % kFanStreams;
        if (!g_fan.used[i]) { cudaStreamWaitEvent(g_fan.st[i], g_fan.start, 0); g_fan.used[i] = true; }
        st = g_fan.st[i];
    }
    return st;
}

int launch_tensor(const Args& a, int op, int cls, const unsigned int* list, unsigned long long n, unsigned long long need, int ext, void* stream) {
    cudaStream_t st = (cudaStream_t)fan_stream(stream);
    const bool fma = a.prm.use_fma != 0;
    if (op == kTRestrict) return fma ? launch_tensor_t<kTRestrict, true>(a, cls, list, n, need, ext, st) : launch_tensor_t<kTRestrict, false>(a, cls, list, n, need, ext, st);
    return fma ? launch_tensor_t<kTSubdivide, true>(a, cls, list, n, need, ext, st) : launch_tensor_t<kTSubdivide, false>(a, cls, list, n, need, ext, st);
}

int launch_scalar(const Args& a, const Segments& s, void* stream) {
    unsigned long long total = s.child_bound;
    for (int c = 0; c < kClasses; ++c) total += s.n[c];
    if (!total) return 0;
    f2_scalar_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a, s);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "f2_scalar_kernel launch");
}

// ---- reporting order: stable LSD radix sort of the record indices by path_lo, path_hi, system (CUB) ----
__global__ void f2_sort_keys(const ResultRec<2>* res, const unsigned int* idx, unsigned long long n, int which,
                             unsigned long long* k64, unsigned int* k32, unsigned int* iota) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int r = idx ? idx[i] : (unsigned int)i;
    if (iota) iota[i] = (unsigned int)i;
    if (which == 0) k64[i] = res[r].path_lo;
    else if (which == 1) k64[i] = res[r].path_hi;
    else k32[i] = (unsigned int)res[r].system;
}

__global__ void f2_pack_sorted(const ResultRec<2>* res, const unsigned int* order, unsigned long long n, yr_sorted_rec* out) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int r = order[i];
    yr_sorted_rec o;
    o.index = r; o.system = res[r].system; o.flags = res[r].flags; o.collector = res[r].collector;
    o.root[0] = res[r].root[0]; o.root[1] = res[r].root[1];
    out[i] = o;
}

int remap_links(void* results, unsigned long long n_results, void* nodes, unsigned long long n_nodes, const RankBases& rb, void* stream) {
    const unsigned long long n = n_results > n_nodes ? n_results : n_nodes;
    if (!n) return 0;
    remap_links_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(static_cast<ResultRec<2>*>(results), n_results,
                                                                                     static_cast<NodeRec<2>*>(nodes), n_nodes, rb);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail(e, "remap_links_kernel launch");
}

void* sort_results(const void* results, unsigned long long n, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const ResultRec<2>* res = static_cast<const ResultRec<2>*>(results);
    const size_t n_ = (size_t)n;
    unsigned long long* k64a = (unsigned long long*)dev_alloc(n_ * 8, stream); unsigned long long* k64b = (unsigned long long*)dev_alloc(n_ * 8, stream);
    unsigned int* k32a = (unsigned int*)dev_alloc(n_ * 4, stream); unsigned int* k32b = (unsigned int*)dev_alloc(n_ * 4, stream);
    unsigned int* ia = (unsigned int*)dev_alloc(n_ * 4, stream); unsigned int* ib = (unsigned int*)dev_alloc(n_ * 4, stream);
    void* result = nullptr;
    if (k64a && k64b && k32a && k32b && ia && ib) {
        const unsigned grid = (unsigned)((n + 255) / 256);
        size_t tb64 = 0, tb32 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb64, k64a, k64b, ia, ib, (int)n, 0, 64, st);
        cub::DeviceRadixSort::SortPairs(nullptr, tb32, k32a, k32b, ia, ib, (int)n, 0, 32, st);
        void* tmp = dev_alloc(tb64 > tb32 ? tb64 : tb32, stream);
        if (tmp) {
            f2_sort_keys<<<grid, 256, 0, st>>>(res, nullptr, n, 0, k64a, nullptr, ia);
            cub::DeviceRadixSort::SortPairs(tmp, tb64, k64a, k64b, ia, ib, (int)n, 0, 64, st);          // by path_lo -> ib
            f2_sort_keys<<<grid, 256, 0, st>>>(res, ib, n, 1, k64a, nullptr, nullptr);
            cub::DeviceRadixSort::SortPairs(tmp, tb64, k64a, k64b, ib, ia, (int)n, 0, 64, st);          // by path_hi -> ia
            f2_sort_keys<<<grid, 256, 0, st>>>(res, ia, n, 2, nullptr, k32a, nullptr);
            cub::DeviceRadixSort::SortPairs(tmp, tb32, k32a, k32b, ia, ib, (int)n, 0, 32, st);          // by system  -> ib
            dev_free(tmp, stream);
            // pack (index, system, root) in reporting order
            yr_sorted_rec* packed = (yr_sorted_rec*)dev_alloc(n_ * sizeof(yr_sorted_rec), stream);
            if (packed) {
                f2_pack_sorted<<<grid, 256, 0, st>>>(res, ib, n, packed);
                if (cudaGetLastError() == cudaSuccess) result = packed; else dev_free(packed, stream);
            }
        }
    }
    dev_free(k64a, stream); dev_free(k64b, stream); dev_free(k32a, stream); dev_free(k32b, stream); dev_free(ia, stream); dev_free(ib, stream);
    return result;
}

double host_ms() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }

double trace_sync_ms(void* stream) {
    static double last = 0.0;
    cudaStreamSynchronize((cudaStream_t)stream);
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    const double now = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6, d = now - last;
    last = now;
    return d;
}

void timer_reset(int enable) { g_timer.enabled = enable; g_timer.used[0] = g_timer.used[1] = 0; }
void timer_mark(int which, int begin, void* stream) {
    if (!g_timer.enabled) return;
    (void)begin;
    auto& v = g_timer.ev[which];
    if (g_timer.used[which] == v.size()) { cudaEvent_t e; cudaEventCreate(&e); v.push_back(e); }
    cudaEventRecord(v[g_timer.used[which]++], (cudaStream_t)stream);
}
void timer_collect(double* tensor_ms, double* scalar_ms) {
    double out[2] = {0.0, 0.0};
    if (g_timer.enabled)
        for (int w = 0; w < 2; ++w)
            for (size_t i = 0; i + 1 < g_timer.used[w]; i += 2) {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, g_timer.ev[w][i], g_timer.ev[w][i + 1]) == cudaSuccess) out[w] += ms;
            }
    *tensor_ms = out[0]; *scalar_ms = out[1];
}

}  // namespace yr_f2_backend

extern "C" {

int yr_frontier_solve(const yr_frontier_desc* d, yr_frontier_out* out, void* stream) {
    if (!d || !out) { yr_set_error("null descriptor"); return -1; }
    if (d->dim == 3 || d->dim == 4) return yr_fd_solve(d, out, stream);           // yr_fd_kernels.cu
    return yr_f2_host::solve(d, out, stream);
}

int yr_frontier_release(yr_frontier_out* out, void* stream) { return yr_f2_host::release(out, stream); }
#include "yr_f2_abi.inc"
int yr_frontier_fits(uint64_t max_extent) { return yr_f2_host::fits(max_extent); }
int yr_frontier_fits_dim(int32_t dim, uint64_t max_extent, uint64_t max_tensor) {
    if (dim == 2) return yr_f2_host::fits(max_extent);
    return yr_fd_fits(dim, max_extent, max_tensor);
}
int yr_frontier_trim(void* stream) { yr_fd_trim(stream); return yr_f2_host::trim_workspace(stream); }

int yr_gather_by_system(const void* table, uint64_t n, uint32_t rec_bytes, uint32_t system_off, const uint8_t* sys_flags, uint64_t nsys,
                        uint32_t* out_index, void* out_recs, uint64_t cap, uint64_t* out_n, void* stream) {
    if (!n) return 0;
    if (rec_bytes % 8) { yr_set_error("yr_gather_by_system: records must be a multiple of 8 bytes"); return -1; }
    gather_by_system_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        static_cast<const unsigned char*>(table), n, rec_bytes, system_off, sys_flags, nsys, out_index, static_cast<unsigned char*>(out_recs), cap,
        reinterpret_cast<unsigned long long*>(out_n));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { yr_set_error(cudaGetErrorString(e)); return -1; }
    return 0;
}

int yr_fetch(void* host_dst, const void* dev_src, uint64_t bytes, void* stream) {
    if (!bytes) return 0;
    return yr_f2_backend::to_host(host_dst, dev_src, (size_t)bytes, stream);
}

}  // extern "C"
