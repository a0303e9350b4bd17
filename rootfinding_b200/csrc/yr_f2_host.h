// Host side of the 2-D frontier pipeline (yr_f2.h): the round loop behind yr_frontier_solve, written against
// backend hooks (namespace yr_f2_backend) that yr_f2_kernels.cu (CUDA, the product) and tests/emul (pthreads,
// test infrastructure) provide.
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../include/yroots_b200.h"
#include "yr_f2.h"

extern "C" void yr_set_error(const char* msg);       // defined next to yr_last_error (yr_api_impl.h)
extern "C" void yr_set_error_oom(const char* msg);   // ... marks the failure as "out of device memory"

namespace yr_f2_backend {
void* dev_alloc(size_t bytes, void* stream);
void dev_free(void* p, void* stream);
int dev_copy(void* dst, const void* src, size_t bytes, void* stream);          // device -> device
int dev_zero(void* p, size_t bytes, void* stream);
int to_host(void* dst, const void* src, size_t bytes, void* stream);           // synchronises the stream
int launch_tables(double* mats, int* rows, int nt, void* stream);              // 4 matrices / 4 rows_after arrays
int launch_import(const yr::f2::Args& a, unsigned long long n, void* stream);
// boxes list[0 .. n): tensor op `op` with `need` doubles of shared memory per team, extents <= ext
int launch_tensor(const yr::f2::Args& a, int op, int cls, const unsigned int* list, unsigned long long n,
                  unsigned long long need, int ext, void* stream);
// the tensor launches of a round are independent: the backend may spread them over several streams between these
void round_fork(void* stream);
void round_join(void* stream);
void* fan_stream(void* stream);          // the stream the next tensor launch of the round should use
struct Segments { const unsigned int* list[yr::f2::kClasses]; unsigned long long n[yr::f2::kClasses];
                  unsigned long long child_lo, child_bound; };
int launch_scalar(const yr::f2::Args& a, const Segments& s, void* stream);
// optional CUDA-event timing of the tensor / scalar launches of a solve
void timer_reset(int enable);
void timer_mark(int which, int begin, void* stream);      // which: 0 tensor, 1 scalar
void timer_collect(double* tensor_ms, double* scalar_ms);
// order[0..n) = record indices sorted by (system, path_hi, path_lo); returns a device buffer (or null)
void* sort_results(const void* results, unsigned long long n, void* stream);
// record links of several ranks' tables (rank << shift | local index) -> positions in their concatenation
struct RankBases { unsigned long long res[16], node[16]; int shift; };
int remap_links(void* results, unsigned long long n_results, void* nodes, unsigned long long n_nodes, const RankBases& rb, void* stream);
double trace_sync_ms(void* stream);      // YR_F2_TRACE: synchronise and return host milliseconds since the previous call
double host_ms();                        // host clock, milliseconds
}  // namespace yr_f2_backend

namespace yr_f2_host {
using namespace yr;
using namespace yr::f2;

extern "C" void yr_set_error_no_axis(void);          // ... as the reference's IndexError (error code 3)
inline int fail(const char* msg) { yr_set_error(msg); return -1; }
inline int fail_no_axis() { yr_set_error_no_axis(); return -1; }

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    // grow to at least `bytes`, keeping the first `keep` bytes
    int reserve(size_t bytes, size_t keep, void* stream, double slack = 1.5) {
        if (bytes <= cap) return 0;
        size_t want = (size_t)(bytes * slack) + 256;
        void* q = yr_f2_backend::dev_alloc(want, stream);
        if (!q) { yr_set_error_oom("device allocation failed (frontier buffers)"); return -1; }
        if (p && keep) { if (yr_f2_backend::dev_copy(q, p, keep, stream)) return -1; }
        if (p) yr_f2_backend::dev_free(p, stream);
        p = q; cap = want;
        return 0;
    }
    void release(void* stream) { if (p) yr_f2_backend::dev_free(p, stream); p = nullptr; cap = 0; }
};

struct Workspace { DevBuf recs, work, arena[2], lists[2], tabs, ctlbuf, results, nodes; int tabs_nt = 0; };
inline Workspace& workspace() { static Workspace w; return w; }
inline int trim_workspace(void* stream) {
    Workspace& w = workspace();
    w.recs.release(stream); w.work.release(stream); w.arena[0].release(stream); w.arena[1].release(stream);
    w.lists[0].release(stream); w.lists[1].release(stream); w.tabs.release(stream); w.ctlbuf.release(stream);
    w.results.release(stream); w.nodes.release(stream);
    w.tabs_nt = 0;
    return 0;
}

inline int fits(unsigned long long max_extent) {
    const int n = (int)(max_extent < 4 ? 4 : max_extent);
    if (max_extent > 4096) return 0;
    const int T = even_up(n * odd_ld(n));
    return need_subdivide(T) + 2ull * panel_doubles(n) <= kMaxNeed && need_restrict(T, n, n) <= kMaxNeed;
}

inline int release(yr_frontier_out* out, void* stream) {
    if (!out) return 0;
    if (out->results) yr_f2_backend::dev_free(out->results, stream);
    if (out->nodes) yr_f2_backend::dev_free(out->nodes, stream);
    if (out->order) yr_f2_backend::dev_free(out->order, stream);
    out->results = nullptr; out->nodes = nullptr; out->order = nullptr;
    return 0;
}

// Result / node records of several processes' frontiers, concatenated by the caller in rank order, become ONE library-owned
// output as if a single solve had produced them: links remapped to positions in the concatenation, reporting order
// computed by the same device sort.  (Box-level multi-GPU: rank 0 receives the other ranks' records over NCCL and goes on
// with the ordinary single-solve post-pass; rootfinding_b200/distributed.py.)
inline int adopt_records(const void* results, unsigned long long n_results, const void* nodes, unsigned long long n_nodes,
                         const uint64_t* res_base, const uint64_t* node_base, int nranks, int rank_shift, yr_frontier_out* out, void* stream) {
    namespace be = yr_f2_backend;
    memset(out, 0, sizeof(*out));
    if (nranks < 1 || nranks > 16) return fail("yr_frontier_adopt: 1..16 ranks");
    be::RankBases rb;
    for (int r = 0; r < 16; ++r) { rb.res[r] = r < nranks ? res_base[r] : 0ull; rb.node[r] = r < nranks ? node_base[r] : 0ull; }
    rb.shift = rank_shift;
    if (n_results) {
        out->results = be::dev_alloc((size_t)n_results * sizeof(ResultRec<2>), stream);
        if (!out->results || be::dev_copy(out->results, results, (size_t)n_results * sizeof(ResultRec<2>), stream)) { release(out, stream); yr_set_error_oom("device allocation failed (results)"); return -1; }
    }
    if (n_nodes) {
        out->nodes = be::dev_alloc((size_t)n_nodes * sizeof(NodeRec<2>), stream);
        if (!out->nodes || be::dev_copy(out->nodes, nodes, (size_t)n_nodes * sizeof(NodeRec<2>), stream)) { release(out, stream); yr_set_error_oom("device allocation failed (nodes)"); return -1; }
    }
    if (be::remap_links(out->results, n_results, out->nodes, n_nodes, rb, stream)) { release(out, stream); return -1; }
    out->order = n_results ? be::sort_results(out->results, n_results, stream) : nullptr;
    if (n_results && !out->order) { release(out, stream); yr_set_error_oom("device allocation failed (order)"); return -1; }
    out->n_results = n_results; out->n_nodes = n_nodes;
    return 0;
}

// ---- a solve as a resumable object ----------------------------------------------------------------------------
// yr_frontier_solve = begin + advance(until empty) + finish.  The pieces are exported on their own
// (yr_frontier_begin / _advance / _export / _import / _finish) so that a host can stop between rounds, move queued
// boxes to another process' frontier (multi-GPU load rebalancing: rootfinding_b200/distributed.py) and go on.
struct Frontier {
    bool active = false;
    Args a; Ctl h;
    unsigned long long n0 = 0, n_boxes = 0, n_results = 0, n_nodes = 0, rounds = 0, launches = 0;
    unsigned long long result_base0 = 0, node_base0 = 0;
    unsigned long long in_used = 0;             // doubles of the current input arena in use (valid once own_input)
    const double* data_in = nullptr;
    bool own_input = false;                     // data_in is one of our arenas (false: the caller's level-0 pool)
    int cur = 0, out_arena = 0, in_arena = -1;
    size_t cap_boxes = 0, cap_results = 0, cap_nodes = 0, per_list[2] = {0, 0};
    int time_kernels = 0, nt = 0;
    bool trace = false, phases = false;
};
inline Frontier& frontier() { static Frontier f; return f; }

// Exchange buffer of yr_frontier_export / _import (device memory): XHeader, BoxRec<2>[n], Work[n], double data[].
// Work::off of an exported box is relative to the data section.
struct XHeader { unsigned long long magic, n_boxes, data_doubles, reserved; };
constexpr unsigned long long kXMagic = 0x5952463258ull;       // "YRF2X"
YR_HD size_t xbuf_rec_off() { return sizeof(XHeader); }
YR_HD size_t xbuf_work_off(unsigned long long n) { return sizeof(XHeader) + (size_t)n * sizeof(BoxRec<2>); }
YR_HD size_t xbuf_data_off(unsigned long long n) { return xbuf_work_off(n) + (size_t)n * sizeof(Work); }
struct XTake { unsigned long long first[kOps][kClasses], count[kOps][kClasses]; };

}  // namespace yr_f2_host

namespace yr_f2_backend {
int to_device(void* dst, const void* src, size_t bytes, void* stream);          // host -> device, synchronises
// pack the queued boxes named by `take` (positions first .. first + count of the pending lists) into `xbuf`
int launch_export(const yr::f2::Args& a, const yr_f2_host::XTake& take, unsigned long long n, unsigned char* xbuf, const double* data_in, void* stream);
// adopt n boxes of `xbuf` as slots slot_base ..: records (offsets shifted by data_base) and queue entries
int launch_import_x(const yr::f2::Args& a, unsigned long long n, const unsigned char* xbuf, unsigned long long slot_base,
                    unsigned long long data_base, void* stream);
}  // namespace yr_f2_backend

namespace yr_f2_host {

// device code shared by both backends --------------------------------------------------------------------------
// one exported box: records + both tensors into the exchange buffer at data offset `at` (lanes copy in parallel; the
// caller synchronises them and then lets one lane call export_fix_offsets)
YR_HD void export_one(const Args& a, unsigned int b, unsigned long long j, unsigned long long n, unsigned char* xbuf,
                      const double* data_in, int lane, int lanes, unsigned long long at) {
    const BoxRec<2>& rec = a.recs[b];
    const Work& wk = a.work[b];
    const int sz0 = tensor_doubles(rec.shape[0][0], wk.ld[0]), sz1 = tensor_doubles(rec.shape[1][0], wk.ld[1]);
    double* data = reinterpret_cast<double*>(xbuf + xbuf_data_off(n));
    for (int e = lane; e < sz0; e += lanes) data[at + e] = data_in[wk.off[0] + e];
    for (int e = lane; e < sz1; e += lanes) data[at + sz0 + e] = data_in[wk.off[1] + e];
    unsigned long long* rdst = reinterpret_cast<unsigned long long*>(xbuf + xbuf_rec_off() + (size_t)j * sizeof(BoxRec<2>));
    const unsigned long long* rsrc = reinterpret_cast<const unsigned long long*>(&rec);
    for (int e = lane; e < (int)(sizeof(BoxRec<2>) / 8); e += lanes) rdst[e] = rsrc[e];
    unsigned long long* wd = reinterpret_cast<unsigned long long*>(reinterpret_cast<Work*>(xbuf + xbuf_work_off(n)) + j);
    const unsigned long long* wsrc = reinterpret_cast<const unsigned long long*>(&wk);
    for (int e = lane; e < (int)(sizeof(Work) / 8); e += lanes) wd[e] = wsrc[e];
}
YR_HD int export_doubles(const Args& a, unsigned int b, int& sz0) {
    const BoxRec<2>& rec = a.recs[b];
    const Work& wk = a.work[b];
    sz0 = tensor_doubles(rec.shape[0][0], wk.ld[0]);
    return sz0 + tensor_doubles(rec.shape[1][0], wk.ld[1]);
}
YR_HD void export_fix_offsets(unsigned long long j, unsigned long long n, unsigned char* xbuf, unsigned long long at, int sz0) {
    Work* w = reinterpret_cast<Work*>(xbuf + xbuf_work_off(n)) + j;
    w->off[0] = at; w->off[1] = at + (unsigned long long)sz0;
}
// queue position j of an export -> the box it names
YR_HD unsigned int export_box_of(const Args& a, const XTake& take, unsigned long long j) {
    for (int o = 0; o < kOps; ++o)
        for (int c = 0; c < kClasses; ++c) {
            if (j < take.count[o][c]) return a.lists[o][c][take.first[o][c] + j];
            j -= take.count[o][c];
        }
    return 0u;
}
// one imported box: records into slot `slot`, decision for its pending tensor op (thread-serial)
template <class At>
YR_HD void import_one(const Args& a, unsigned long long j, unsigned long long n, const unsigned char* xbuf, unsigned int slot,
                      unsigned long long data_base, Decision& dc, At at) {
    const BoxRec<2>* rsrc = reinterpret_cast<const BoxRec<2>*>(xbuf + xbuf_rec_off()) + j;
    const Work* wsrc = reinterpret_cast<const Work*>(xbuf + xbuf_work_off(n)) + j;
    a.recs[slot] = *rsrc;
    Work w = *wsrc;
    w.off[0] += data_base; w.off[1] += data_base;
    a.work[slot] = w;
    const BoxRec<2>& rec = a.recs[slot];
    int T, n0m, n1m; unsigned long long total;
    box_sizes(rec, w, T, n0m, n1m, total);
    dc.op = w.op;
    dc.ext = (unsigned long long)(n0m > n1m ? n0m : n1m);
    if (w.op == kTRestrict) { pick_team(need_restrict_w(T, n0m, n1m), need_restrict(T, n0m, n1m), a.coarse, a.pol, 0, dc.need, dc.cls); dc.bound = total; }
    else if (w.op == kTSubdivide) { pick_team(need_subdivide_w(T), need_subdivide(T), a.coarse, a.pol, 1, dc.need, dc.cls); dc.bound = total << w.nsplit; }
    (void)at;
}

inline void carve(Frontier& f, DevBuf& buf, size_t per_list) {
    unsigned int* base = static_cast<unsigned int*>(buf.p);
    for (int o = 0; o < kOps; ++o) for (int c = 0; c < kClasses; ++c) f.a.lists[o][c] = base + (size_t)(o * kClasses + c) * per_list;
}
inline void bind_tables(Frontier& f) {
    Workspace& ws = workspace();
    f.a.recs = static_cast<BoxRec<2>*>(ws.recs.p); f.a.work = static_cast<Work*>(ws.work.p); f.a.cap_boxes = f.cap_boxes;
    f.a.results = static_cast<ResultRec<2>*>(ws.results.p); f.a.cap_results = f.cap_results;
    f.a.nodes = static_cast<NodeRec<2>*>(ws.nodes.p); f.a.cap_nodes = f.cap_nodes;
}
inline unsigned long long pending(const Ctl& h) {
    unsigned long long t = 0;
    for (int o = 0; o < kOps; ++o) for (int c = 0; c < kClasses; ++c) t += h.n_list[o][c];
    return t;
}

#define F2_TRY(expr) do { if (expr) { f.active = false; return -1; } } while (0)

inline int begin(const yr_frontier_desc* d, void* stream) {
    namespace be = yr_f2_backend;
    Frontier& f = frontier();
    if (f.active) return fail("yr_frontier_begin: a frontier solve is already in progress (one per process)");
    if (d->dim != 2) return fail("yr_frontier_solve: dim must be 2");
    if (d->prm.exact) return fail("yr_frontier_solve: exact mode is served by yr_process_level");
    if (d->prm.trim_policy) return fail("yr_frontier_solve: the alternative trim policies are served by yr_process_level");
    const unsigned long long n0 = d->n_in;
    if (n0 > 0x3FFFFFFFull) return fail("yr_frontier_solve: too many level-0 boxes");
    const int nt = (int)(d->max_extent < 4 ? 4 : d->max_extent);
    if (!fits(d->max_extent)) return fail("yr_frontier_solve: tensors too large for the shared-memory pipeline");
    f = Frontier();
    f.active = true; f.n0 = n0; f.nt = nt; f.time_kernels = d->time_kernels;
    be::timer_reset(d->time_kernels);
    f.trace = getenv("YR_F2_TRACE") != nullptr;
    f.phases = getenv("YR_F2_PHASES") != nullptr;       // coarse host-clock phase summary (synchronises)
    if (f.phases) be::trace_sync_ms(stream);

    // Box tables, arenas, queues, matrix tables and the control block persist between solves (grow-only, released
    // by yr_frontier_trim); result / node records accumulate in persistent buffers sized by the (loose) per-round
    // bounds, the caller gets exactly sized copies at the end.
    Workspace& ws = workspace();
    // level-independent subdivision matrices
    const size_t P = (size_t)panel_doubles(nt);
    F2_TRY(ws.tabs.reserve(4 * P * 8 + 4 * (size_t)nt * 4 + 64, 0, stream, 1.0));
    double* mats = static_cast<double*>(ws.tabs.p);
    int* rows = reinterpret_cast<int*>(mats + 4 * P);
    if (ws.tabs_nt != nt) { F2_TRY(be::launch_tables(mats, rows, nt, stream)); ws.tabs_nt = nt; }

    F2_TRY(ws.ctlbuf.reserve(sizeof(Ctl), 0, stream, 1.0));
    F2_TRY(be::dev_zero(ws.ctlbuf.p, sizeof(Ctl), stream));
    f.cap_boxes = (size_t)(n0 * 8 + 4096);
    F2_TRY(ws.recs.reserve(f.cap_boxes * sizeof(BoxRec<2>), 0, stream, 1.0));
    F2_TRY(ws.work.reserve(f.cap_boxes * sizeof(Work), 0, stream, 1.0));
    { const size_t have = ws.recs.cap / sizeof(BoxRec<2>) < ws.work.cap / sizeof(Work) ? ws.recs.cap / sizeof(BoxRec<2>) : ws.work.cap / sizeof(Work); if (have > f.cap_boxes) f.cap_boxes = have; }
    if (n0) F2_TRY(be::dev_copy(ws.recs.p, d->recs_in, n0 * sizeof(BoxRec<2>), stream));
    f.cap_results = (size_t)(n0 * 4 + 1024); f.cap_nodes = f.cap_results;
    F2_TRY(ws.results.reserve(f.cap_results * sizeof(ResultRec<2>), 0, stream, 1.0));
    F2_TRY(ws.nodes.reserve(f.cap_nodes * sizeof(NodeRec<2>), 0, stream, 1.0));
    if (ws.results.cap / sizeof(ResultRec<2>) > f.cap_results) f.cap_results = ws.results.cap / sizeof(ResultRec<2>);
    if (ws.nodes.cap / sizeof(NodeRec<2>) > f.cap_nodes) f.cap_nodes = ws.nodes.cap / sizeof(NodeRec<2>);

    Args& a = f.a;
    memset(&a, 0, sizeof(a));
    a.ctl = static_cast<Ctl*>(ws.ctlbuf.p);
    a.tab.nt = nt;
    for (int m = 0; m < 2; ++m) for (int h = 0; h < 2; ++h) { a.tab.mat[m][h] = mats + (size_t)(2 * m + h) * P; a.tab.rows_after[m][h] = rows + (size_t)(2 * m + h) * nt; }
    memcpy(&a.prm, &d->prm, sizeof(a.prm));
    a.result_base = d->result_base; a.node_base = d->node_base;
    f.result_base0 = d->result_base; f.node_base0 = d->node_base;
    {
        // team policy: warp teams up to these needs (doubles), one CTA per box above; overridable for experiments
        auto env_ull = [](const char* name, unsigned long long dflt) { const char* v = getenv(name); return v ? strtoull(v, nullptr, 10) : dflt; };
        a.pol.warp_max[0] = env_ull("YR_F2_WARP_MAX_R", kWarp32Max);
        a.pol.warp_max[1] = env_ull("YR_F2_WARP_MAX_S", kWarp32Max);
        a.pol.cta128_max = env_ull("YR_F2_CTA128_MAX", 7168ull);
        for (int o = 0; o < 2; ++o) if (a.pol.warp_max[o] > kWarp32Max) a.pol.warp_max[o] = kWarp32Max;
        if (a.pol.cta128_max > 7168ull) a.pol.cta128_max = 7168ull;
    }
    memset(&f.h, 0, sizeof(f.h));
    // round -1: adopt the level-0 boxes (they arrive packed, from yr_init_boxes) and queue their first pass
    f.cur = 0;
    f.per_list[0] = (size_t)(n0 ? n0 : 1);
    F2_TRY(ws.lists[0].reserve((size_t)kOps * kClasses * f.per_list[0] * 4 + 64, 0, stream, 1.25));
    carve(f, ws.lists[0], f.per_list[0]);
    bind_tables(f);
    a.data_in = d->data_in;
    f.data_in = d->data_in; f.own_input = false; f.in_arena = -1; f.in_used = 0;
    f.n_boxes = n0; f.launches = 2;
    if (n0) {
        F2_TRY(be::launch_import(a, n0, stream));
        F2_TRY(be::to_host(&f.h, ws.ctlbuf.p, sizeof(Ctl), stream));
    }
    f.out_arena = 0;
    return 0;
}

// runs rounds until the queues are empty or `max_rounds` rounds have run (0: no limit)
inline int advance(unsigned long long max_rounds, void* stream) {
    namespace be = yr_f2_backend;
    Frontier& f = frontier();
    if (!f.active) return fail("yr_frontier_advance: no frontier solve in progress");
    Workspace& ws = workspace();
    Args& a = f.a; Ctl& h = f.h;
    unsigned int* cur_lists[kOps][kClasses];
    unsigned long long done_rounds = 0;
    for (;;) {
        unsigned long long n_res = 0, n_sub = 0;
        for (int c = 0; c < kClasses; ++c) { n_res += h.n_list[0][c]; n_sub += h.n_list[1][c]; }
        if (n_res + n_sub == 0) break;
        if (max_rounds && done_rounds >= max_rounds) break;
        ++done_rounds;
        const double t_round0 = f.phases ? be::host_ms() : 0.0;
        if (++f.rounds > 1000000ull) { f.active = false; return fail("yr_frontier_solve: round loop did not terminate"); }
        const unsigned long long produced = n_res + 4 * n_sub;            // boxes the scalar step of this round may see
        for (int o = 0; o < kOps; ++o) for (int c = 0; c < kClasses; ++c) cur_lists[o][c] = a.lists[o][c];
        // capacities for everything this round can append
        if (f.n_boxes + 4 * n_sub > f.cap_boxes) {
            const size_t want = (size_t)((f.n_boxes + 4 * n_sub) * 2) + 4096;
            F2_TRY(ws.recs.reserve(want * sizeof(BoxRec<2>), f.n_boxes * sizeof(BoxRec<2>), stream, 1.0));
            F2_TRY(ws.work.reserve(want * sizeof(Work), f.n_boxes * sizeof(Work), stream, 1.0));
            f.cap_boxes = want;
        }
        if (f.n_results + produced > f.cap_results) {
            const size_t want = (size_t)((f.n_results + produced) * 2) + 1024;
            F2_TRY(ws.results.reserve(want * sizeof(ResultRec<2>), f.n_results * sizeof(ResultRec<2>), stream, 1.0));
            f.cap_results = want;
        }
        if (f.n_nodes + produced > f.cap_nodes) {
            const size_t want = (size_t)((f.n_nodes + produced) * 2) + 1024;
            F2_TRY(ws.nodes.reserve(want * sizeof(NodeRec<2>), f.n_nodes * sizeof(NodeRec<2>), stream, 1.0));
            f.cap_nodes = want;
        }
        F2_TRY(ws.arena[f.out_arena].reserve((size_t)(h.out_bound + 16) * 8, 0, stream, 1.25));
        f.per_list[f.cur ^ 1] = (size_t)produced;
        F2_TRY(ws.lists[f.cur ^ 1].reserve((size_t)kOps * kClasses * f.per_list[f.cur ^ 1] * 4 + 64, 0, stream, 1.25));
        carve(f, ws.lists[f.cur ^ 1], f.per_list[f.cur ^ 1]);
        bind_tables(f);
        a.results = static_cast<ResultRec<2>*>(ws.results.p) + f.n_results; a.cap_results = f.cap_results - f.n_results;
        a.nodes = static_cast<NodeRec<2>*>(ws.nodes.p) + f.n_nodes; a.cap_nodes = f.cap_nodes - f.n_nodes;
        a.result_base = f.result_base0 + f.n_results; a.node_base = f.node_base0 + f.n_nodes;
        a.coarse = (produced < 8192ull) ? 1 : 0;
        a.data_in = f.data_in; a.data_out = static_cast<double*>(ws.arena[f.out_arena].p); a.cap_data_out = ws.arena[f.out_arena].cap / 8;
        // per-round part of the control block: queues, bounds, arena cursor; slot counters restart at their totals
        F2_TRY(be::dev_zero(ws.ctlbuf.p, offsetof(Ctl, n_boxes), stream));
        F2_TRY(be::dev_zero(&a.ctl->n_results, 2 * sizeof(unsigned long long), stream));
        if (f.trace) fprintf(stderr, "f2 round %3llu: host %.3f ms |", f.rounds, be::trace_sync_ms(stream));
        be::timer_mark(0, 1, stream);
        be::round_fork(stream);
        for (int o = kOps - 1; o >= 0; --o)
            for (int c = 0; c < kClasses; ++c) {
                if (!h.n_list[o][c]) continue;
                if (h.need_max[o][c] + (o ? 2ull * panel_doubles((int)(h.ext_max[c] < 4 ? 4 : h.ext_max[c])) : 0ull) > kMaxNeed) { f.active = false; return fail("yr_frontier_solve: a box exceeds the shared-memory budget"); }
                F2_TRY(be::launch_tensor(a, o == 0 ? (int)kTRestrict : (int)kTSubdivide, c, cur_lists[o][c], h.n_list[o][c],
                                         h.need_max[o][c], (int)(o ? h.ext_max[c] : 0), stream));
                ++f.launches;
                if (f.trace) fprintf(stderr, " %s%d n=%llu need=%llu ext=%llu %.3f ms |", o ? "S" : "R", c, h.n_list[o][c], h.need_max[o][c], o ? h.ext_max[c] : 0ull, be::trace_sync_ms(stream));
            }
        be::round_join(stream);
        be::timer_mark(0, 0, stream);
        be::Segments sg;
        for (int c = 0; c < kClasses; ++c) { sg.list[c] = cur_lists[0][c]; sg.n[c] = h.n_list[0][c]; }
        sg.child_lo = f.n_boxes; sg.child_bound = 4 * n_sub;
        be::timer_mark(1, 1, stream);
        F2_TRY(be::launch_scalar(a, sg, stream));
        be::timer_mark(1, 0, stream);
        ++f.launches;
        if (f.trace) fprintf(stderr, " scalar %.3f ms\n", be::trace_sync_ms(stream));
        const double t_issue = f.phases ? be::host_ms() : 0.0;
        F2_TRY(be::to_host(&h, ws.ctlbuf.p, sizeof(Ctl), stream));
        if (f.phases) { const double t_done = be::host_ms(); fprintf(stderr, " r%llu:%.2f+%.2f", f.rounds, t_issue - t_round0, t_done - t_issue); }
        if (h.overflow) { f.active = false; return fail("yr_frontier_solve: internal buffer overflow"); }
        if (h.bad_mid) { f.active = false; return fail("yr_frontier_solve: unexpected split point (nextTransformPoints)"); }
        if (h.no_axis) { f.active = false; return fail_no_axis(); }
        f.n_boxes = h.n_boxes; f.n_results += h.n_results; f.n_nodes += h.n_nodes;
        f.data_in = a.data_out; f.own_input = true; f.in_arena = f.out_arena; f.in_used = h.data_used;
        f.out_arena ^= 1; f.cur ^= 1;
    }
    return 0;
}

inline int finish(yr_frontier_out* out, void* stream) {
    namespace be = yr_f2_backend;
    Frontier& f = frontier();
    memset(out, 0, sizeof(*out));
    if (!f.active) return fail("yr_frontier_finish: no frontier solve in progress");
    Workspace& ws = workspace();
    f.active = false;
    const unsigned long long n_results = f.n_results, n_nodes = f.n_nodes;
    double ph2 = 0.0;
    out->order = n_results ? be::sort_results(ws.results.p, n_results, stream) : nullptr;
    if (f.phases) { ph2 = be::trace_sync_ms(stream); fprintf(stderr, "f2 phases: sort %.2f ms (%llu rounds, %llu launches)\n", ph2, f.rounds, f.launches); }
    if (n_results) {
        out->results = be::dev_alloc((size_t)n_results * sizeof(ResultRec<2>), stream);
        if (!out->results || be::dev_copy(out->results, ws.results.p, (size_t)n_results * sizeof(ResultRec<2>), stream)) { release(out, stream); yr_set_error_oom("device allocation failed (results)"); return -1; }
    }
    if (n_nodes) {
        out->nodes = be::dev_alloc((size_t)n_nodes * sizeof(NodeRec<2>), stream);
        if (!out->nodes || be::dev_copy(out->nodes, ws.nodes.p, (size_t)n_nodes * sizeof(NodeRec<2>), stream)) { release(out, stream); yr_set_error_oom("device allocation failed (nodes)"); return -1; }
    }
    out->n_results = n_results; out->n_nodes = n_nodes;
    out->rounds = f.rounds; out->launches = f.launches; out->slots = f.n_boxes;
    yr_counters& c = out->counters;
    const Ctl& h = f.h;
    c.n_results = n_results; c.n_nodes = n_nodes; c.boxes = h.boxes; c.zooms = h.zooms; c.subdivisions = h.subdivisions;
    c.discard_const = h.dconst; c.discard_quad = h.dquad; c.discard_zoom = h.dzoom; c.entry_coeffs = h.entry_coeffs;
    c.restrict_macs = h.restrict_macs; c.max_level = h.max_level;
    be::timer_collect(&out->tensor_ms, &out->scalar_ms);
    return 0;
}

inline int solve(const yr_frontier_desc* d, yr_frontier_out* out, void* stream) {
    memset(out, 0, sizeof(*out));
    if (d->n_in == 0) return 0;
    if (begin(d, stream)) return -1;
    if (advance(0, stream)) return -1;
    return finish(out, stream);
}

// ---- moving queued boxes between frontiers (multi-GPU load rebalancing) ------------------------------------
inline unsigned long long xbuf_bytes_bound(unsigned long long n_take) {
    const Frontier& f = frontier();
    // every exported tensor lives in the input arena: its used size bounds the data section
    return (unsigned long long)xbuf_data_off(n_take) + (f.in_used + 2 * n_take + 16) * 8;
}

// Removes up to n_take queued boxes (the tails of the pending lists, every (op, class) in proportion) and packs them
// into xbuf.  *n_taken / *bytes: what was written.
inline int export_boxes(unsigned long long n_take, void* xbuf, unsigned long long cap_bytes, unsigned long long* n_taken,
                        unsigned long long* bytes, void* stream) {
    namespace be = yr_f2_backend;
    Frontier& f = frontier();
    *n_taken = 0; *bytes = 0;
    if (!f.active) return fail("yr_frontier_export: no frontier solve in progress");
    Ctl& h = f.h;
    const unsigned long long total = pending(h);
    if (n_take > total) n_take = total;
    if (!n_take) return 0;
    if (!f.own_input) return fail("yr_frontier_export: run at least one round first (the level-0 pool belongs to the caller)");
    if (cap_bytes < xbuf_bytes_bound(n_take)) return fail("yr_frontier_export: exchange buffer too small (yr_frontier_export_bound)");
    XTake take;
    unsigned long long got = 0;
    for (int o = 0; o < kOps; ++o) for (int c = 0; c < kClasses; ++c) {
        take.count[o][c] = (unsigned long long)((long double)h.n_list[o][c] * n_take / total);
        got += take.count[o][c];
    }
    for (int o = 0; o < kOps && got < n_take; ++o) for (int c = 0; c < kClasses && got < n_take; ++c) {
        const unsigned long long room = h.n_list[o][c] - take.count[o][c];
        const unsigned long long add = room < n_take - got ? room : n_take - got;
        take.count[o][c] += add; got += add;
    }
    for (int o = 0; o < kOps; ++o) for (int c = 0; c < kClasses; ++c) take.first[o][c] = h.n_list[o][c] - take.count[o][c];
    XHeader hdr = {kXMagic, n_take, 0ull, 0ull};
    F2_TRY(be::to_device(xbuf, &hdr, sizeof(hdr), stream));
    F2_TRY(be::launch_export(f.a, take, n_take, static_cast<unsigned char*>(xbuf), f.data_in, stream));
    F2_TRY(be::to_host(&hdr, xbuf, sizeof(hdr), stream));
    for (int o = 0; o < kOps; ++o) for (int c = 0; c < kClasses; ++c) h.n_list[o][c] -= take.count[o][c];
    F2_TRY(be::to_device(&f.a.ctl->n_list[0][0], &h.n_list[0][0], sizeof(h.n_list), stream));
    *n_taken = n_take;
    *bytes = (unsigned long long)xbuf_data_off(n_take) + hdr.data_doubles * 8;
    return 0;
}

// Adopts the boxes of an exchange buffer written by yr_frontier_export (of this or another process).
inline int import_boxes(const void* xbuf, unsigned long long bytes, void* stream) {
    namespace be = yr_f2_backend;
    Frontier& f = frontier();
    if (!f.active) return fail("yr_frontier_import: no frontier solve in progress");
    if (bytes < sizeof(XHeader)) return fail("yr_frontier_import: truncated exchange buffer");
    Workspace& ws = workspace();
    XHeader hdr;
    F2_TRY(be::to_host(&hdr, xbuf, sizeof(hdr), stream));
    if (hdr.magic != kXMagic) return fail("yr_frontier_import: not an exchange buffer");
    const unsigned long long n = hdr.n_boxes, nd = hdr.data_doubles;
    if (!n) return 0;
    if (bytes < xbuf_data_off(n) + nd * 8) return fail("yr_frontier_import: truncated exchange buffer");
    Ctl& h = f.h;
    // the imported tensors join the input arena of the next round
    if (!f.own_input) {
        // nothing has run here yet: the (possibly empty) level-0 pool is the caller's -- move it into an arena of ours
        if (f.n0) return fail("yr_frontier_import: run at least one round before importing into a frontier that has level-0 boxes");
        f.in_arena = f.out_arena ^ 1; f.in_used = 0; f.own_input = true;
    }
    DevBuf& ar = ws.arena[f.in_arena];
    F2_TRY(ar.reserve((size_t)(f.in_used + nd + 16) * 8, (size_t)f.in_used * 8, stream, 1.25));
    f.data_in = static_cast<const double*>(ar.p);
    F2_TRY(be::dev_copy(static_cast<double*>(ar.p) + f.in_used, static_cast<const unsigned char*>(xbuf) + xbuf_data_off(n), (size_t)nd * 8, stream));
    // slots
    if (f.n_boxes + n > f.cap_boxes) {
        const size_t want = (size_t)((f.n_boxes + n) * 2) + 4096;
        F2_TRY(ws.recs.reserve(want * sizeof(BoxRec<2>), f.n_boxes * sizeof(BoxRec<2>), stream, 1.0));
        F2_TRY(ws.work.reserve(want * sizeof(Work), f.n_boxes * sizeof(Work), stream, 1.0));
        f.cap_boxes = want;
    }
    // queues: every pending list must have room for all incoming boxes
    unsigned long long longest = 0;
    for (int o = 0; o < kOps; ++o) for (int c = 0; c < kClasses; ++c) if (h.n_list[o][c] > longest) longest = h.n_list[o][c];
    if (longest + n > f.per_list[f.cur]) {
        const size_t per = (size_t)((longest + n) * 3 / 2 + 64);
        DevBuf fresh;
        F2_TRY(fresh.reserve((size_t)kOps * kClasses * per * 4 + 64, 0, stream, 1.0));
        unsigned int* base = static_cast<unsigned int*>(fresh.p);
        for (int o = 0; o < kOps; ++o) for (int c = 0; c < kClasses; ++c)
            if (h.n_list[o][c]) F2_TRY(be::dev_copy(base + (size_t)(o * kClasses + c) * per, f.a.lists[o][c], (size_t)h.n_list[o][c] * 4, stream));
        ws.lists[f.cur].release(stream);
        ws.lists[f.cur] = fresh;
        f.per_list[f.cur] = per;
        carve(f, ws.lists[f.cur], per);
    }
    bind_tables(f);
    f.a.coarse = 0;
    f.a.data_in = f.data_in;
    F2_TRY(be::launch_import_x(f.a, n, static_cast<const unsigned char*>(xbuf), f.n_boxes, f.in_used, stream));
    Ctl now;
    F2_TRY(be::to_host(&now, ws.ctlbuf.p, sizeof(Ctl), stream));
    memcpy(h.n_list, now.n_list, sizeof(h.n_list)); memcpy(h.need_max, now.need_max, sizeof(h.need_max));
    memcpy(h.ext_max, now.ext_max, sizeof(h.ext_max)); h.out_bound = now.out_bound;
    f.n_boxes += n; f.in_used += nd;
    // the slot counter on the device is what the next round's subdivisions allocate from
    F2_TRY(be::to_device(&f.a.ctl->n_boxes, &f.n_boxes, sizeof(unsigned long long), stream));
    h.n_boxes = f.n_boxes;
    return 0;
}
// ---- operator-level entry: the frontier pipeline's tensor kernels on caller-supplied boxes -----------------------
// yr_f2_tensor_ops (include/yroots_b200.h).  Everything crosses the boundary through HOST arrays: this is the batched
// form of transformChebToInterval (:864-894) / getSubdivisionIntervals (:1051-1129) for callers and parity tests, not
// the solve path.  The boxes go through exactly the kernels yr_frontier_solve launches (same classes, same teams).
inline int tensor_ops(const yr_f2_ops_desc* d, void* stream) {
    namespace be = yr_f2_backend;
    Frontier& f = frontier();
    if (f.active) return fail("yr_f2_tensor_ops: a frontier solve is in progress");
    const unsigned long long n = d->n;
    if (!n) return 0;
    const bool sub = d->op == (int)kTSubdivide;
    if (!sub && d->op != (int)kTRestrict) return fail("yr_f2_tensor_ops: op must be 1 (restrict) or 2 (subdivide)");
    int ext = 4;
    for (unsigned long long b = 0; b < n; ++b) for (int p = 0; p < 2; ++p) for (int a = 0; a < 2; ++a) {
        const int e = d->shape[(b * 2 + p) * 2 + a];
        if (e < 1) return fail("yr_f2_tensor_ops: extents must be >= 1");
        if (e > ext) ext = e;
    }
    if (!fits((unsigned long long)ext)) return fail("yr_f2_tensor_ops: tensors too large for the shared-memory pipeline");
    yr_frontier_desc fd;
    memset(&fd, 0, sizeof(fd));
    fd.dim = 2; fd.n_in = 0; fd.max_extent = (unsigned long long)ext; fd.prm = d->prm;
    fd.prm.constant_check = 0;          // every output tensor is wanted here (a solve does not write children the check discards)
    if (begin(&fd, stream)) return -1;
    Workspace& ws = workspace();
    Args& a = f.a;
    // ---- host images of the records and of the input arena ----
    std::vector<BoxRec<2>> recs(n);
    std::vector<Work> work(n);
    std::vector<double> arena;
    std::vector<std::vector<unsigned int>> lists((size_t)kClasses);
    unsigned long long need_max[kClasses], ext_max[kClasses], out_bound = 0;
    for (int c = 0; c < kClasses; ++c) need_max[c] = ext_max[c] = 0;
    const int slot = sub ? 1 : 0;
    for (unsigned long long b = 0; b < n; ++b) {
        BoxRec<2>& rec = recs[b]; Work& w = work[b];
        memset(&rec, 0, sizeof(rec)); memset(&w, 0, sizeof(w));
        rec.st.init_unit();
        rec.system = (int32_t)b; rec.parent_node = -1; rec.collector = -1; rec.level = d->level ? d->level[b] : 0;
        for (int p = 0; p < 2; ++p) {
            const int n0 = d->shape[(b * 2 + p) * 2], n1 = d->shape[(b * 2 + p) * 2 + 1];
            rec.shape[p][0] = n0; rec.shape[p][1] = n1; rec.err[p] = d->err[b * 2 + p];
            const int ld = (d->flat || sub) ? odd_ld(n1) : n1;          // (a subdivision only ever sees arena tensors: odd row stride)
            if (arena.size() & 1) arena.push_back(0.0);
            w.off[p] = arena.size(); w.ld[p] = ld;
            const double* src = d->coeffs + d->off[b * 2 + p];
            double sum = 0.0;
            const size_t base = arena.size();
            arena.resize(base + (size_t)tensor_doubles(n0, ld), 0.0);
            for (int i = 0; i < n0; ++i) for (int j = 0; j < n1; ++j) { const double v = src[(size_t)i * n1 + j]; arena[base + (size_t)i * ld + j] = v; sum += fabs(v); }
            w.sumabs[p] = sum; w.tsumabs[p] = sum; w.terr[p] = rec.err[p];
            w.tshape[p][0] = n0; w.tshape[p][1] = n1;
            w.low[p][0] = src[0];
        }
        if (d->iv) for (int ax = 0; ax < 2; ++ax) { rec.st.iv[ax][0] = d->iv[(b * 2 + ax) * 2]; rec.st.iv[ax][1] = d->iv[(b * 2 + ax) * 2 + 1]; }
        if (d->next_mid) for (int ax = 0; ax < 2; ++ax) rec.st.next_mid[ax] = d->next_mid[b * 2 + ax];
        for (int ax = 0; ax < 2; ++ax) { w.cell_iv[ax][0] = rec.st.iv[ax][0]; w.cell_iv[ax][1] = rec.st.iv[ax][1]; }
        w.node_level = rec.level + 1; w.result_index = -1; w.node_index = (int32_t)b;      // children report parent_node = b
        w.trim_valid = 1; w.changed = 1;
        int T, n0m, n1m; unsigned long long total, need; int cls;
        box_sizes(rec, w, T, n0m, n1m, total);
        if (!sub) {
            w.op = kTRestrict;
            for (int ax = 0; ax < 2; ++ax) { w.alpha[ax] = d->alpha[b * 2 + ax]; w.beta[ax] = d->beta[b * 2 + ax]; }
            pick_team(need_restrict_w(T, n0m, n1m), need_restrict(T, n0m, n1m), 0, a.pol, 0, need, cls);
            out_bound += total;
        } else {
            w.op = kTSubdivide;
            int shape[2][2] = {{rec.shape[0][0], rec.shape[0][1]}, {rec.shape[1][0], rec.shape[1][1]}};
            int order[2][2];
            const int k = subdivision_axes<2>(shape, rec.st.iv, w.node_level, order);     // getSubdivisionDims :1013-1049
            w.nsplit = k;
            for (int p = 0; p < 2; ++p) { w.order[p][0] = order[p][0]; w.order[p][1] = (k > 1) ? order[p][1] : order[p][0]; }
            pick_team(need_subdivide_w(T), need_subdivide(T), 0, a.pol, 1, need, cls);
            out_bound += total << k;
        }
        lists[(size_t)cls].push_back((unsigned int)b);
        if (need > need_max[cls]) need_max[cls] = need;
        const unsigned long long e = (unsigned long long)(n0m > n1m ? n0m : n1m);
        if (e > ext_max[cls]) ext_max[cls] = e;
    }
    // ---- device buffers ----
#define F2_TRY2(expr) do { if (expr) { f.active = false; return -1; } } while (0)
    f.cap_boxes = (size_t)(n * 5 + 64);
    F2_TRY2(ws.recs.reserve(f.cap_boxes * sizeof(BoxRec<2>), 0, stream, 1.0));
    F2_TRY2(ws.work.reserve(f.cap_boxes * sizeof(Work), 0, stream, 1.0));
    F2_TRY2(be::to_device(ws.recs.p, recs.data(), n * sizeof(BoxRec<2>), stream));
    F2_TRY2(be::to_device(ws.work.p, work.data(), n * sizeof(Work), stream));
    F2_TRY2(ws.arena[0].reserve((arena.size() + 16) * 8, 0, stream, 1.0));
    F2_TRY2(be::to_device(ws.arena[0].p, arena.data(), arena.size() * 8, stream));
    F2_TRY2(ws.arena[1].reserve((size_t)(out_bound + 16) * 8, 0, stream, 1.0));
    F2_TRY2(ws.lists[0].reserve((size_t)n * 4 + 64, 0, stream, 1.0));
    bind_tables(f);
    a.data_in = static_cast<const double*>(ws.arena[0].p);
    a.data_out = static_cast<double*>(ws.arena[1].p); a.cap_data_out = ws.arena[1].cap / 8;
    unsigned long long nb = n;
    F2_TRY2(be::to_device(&a.ctl->n_boxes, &nb, sizeof(nb), stream));
    unsigned int* dl = static_cast<unsigned int*>(ws.lists[0].p);
    size_t pos = 0;
    for (int c = 0; c < kClasses; ++c) {
        if (lists[(size_t)c].empty()) continue;
        F2_TRY2(be::to_device(dl + pos, lists[(size_t)c].data(), lists[(size_t)c].size() * 4, stream));
        F2_TRY2(be::launch_tensor(a, d->op, c, dl + pos, lists[(size_t)c].size(), need_max[c], (int)(sub ? ext_max[c] : 0), stream));
        pos += lists[(size_t)c].size();
    }
    // ---- read everything back and unpack on the host ----
    Ctl h;
    F2_TRY2(be::to_host(&h, ws.ctlbuf.p, sizeof(Ctl), stream));
    if (h.overflow) { f.active = false; return fail("yr_f2_tensor_ops: internal buffer overflow"); }
    const unsigned long long slots = h.n_boxes;
    recs.resize(slots); work.resize(slots);
    F2_TRY2(be::to_host(recs.data(), ws.recs.p, slots * sizeof(BoxRec<2>), stream));
    F2_TRY2(be::to_host(work.data(), ws.work.p, slots * sizeof(Work), stream));
    std::vector<double> outd((size_t)h.data_used + 2);
    if (h.data_used) F2_TRY2(be::to_host(outd.data(), ws.arena[1].p, (size_t)h.data_used * 8, stream));
#undef F2_TRY2
    f.active = false;
    const unsigned long long first = sub ? n : 0, count = sub ? slots - n : n;
    if (count > d->cap_out) return fail("yr_f2_tensor_ops: output arrays too small (cap_out)");
    unsigned long long used = 0;
    for (unsigned long long q = 0; q < count; ++q) {
        const BoxRec<2>& rec = recs[first + q]; const Work& w = work[first + q];
        d->out_parent[q] = sub ? rec.parent_node : (int32_t)q;
        for (int p = 0; p < 2; ++p) {
            const int n0 = rec.shape[p][0], n1 = rec.shape[p][1];
            d->out_shape[(q * 2 + p) * 2] = n0; d->out_shape[(q * 2 + p) * 2 + 1] = n1;
            d->out_err[q * 2 + p] = rec.err[p]; d->out_sumabs[q * 2 + p] = w.sumabs[p];
            d->out_trim_shape[(q * 2 + p) * 2] = w.tshape[p][0]; d->out_trim_shape[(q * 2 + p) * 2 + 1] = w.tshape[p][1];
            d->out_trim_err[q * 2 + p] = w.terr[p];
            if (used + (unsigned long long)n0 * n1 > d->cap_coeffs) return fail("yr_f2_tensor_ops: output arrays too small (cap_coeffs)");
            d->out_off[q * 2 + p] = used;
            for (int i = 0; i < n0; ++i) for (int j = 0; j < n1; ++j) d->out_coeffs[used + (size_t)i * n1 + j] = outd[(size_t)w.off[p] + (size_t)i * w.ld[p] + j];
            used += (unsigned long long)n0 * n1;
        }
        for (int ax = 0; ax < 2; ++ax) {
            d->out_iv[(q * 2 + ax) * 2] = rec.st.iv[ax][0]; d->out_iv[(q * 2 + ax) * 2 + 1] = rec.st.iv[ax][1];
            d->out_next_mid[q * 2 + ax] = rec.st.next_mid[ax];
        }
    }
    *d->n_out = count;
    return 0;
}

#undef F2_TRY

}  // namespace yr_f2_host
