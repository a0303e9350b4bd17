// Host side of the D-dimensional frontier pipeline (yr_fd.h, D = 3, 4): the round loop of yr_f2_host.h on the D-dimensional
// records.  Same workspace discipline (grow-only buffers per dimension), same control block, same backend allocation hooks.
#pragma once
#include "yr_f2_host.h"
#include "yr_fd.h"

namespace yr_fd_backend {
struct Segments { const unsigned int* list[yr::f2::kClasses]; unsigned long long n[yr::f2::kClasses]; unsigned long long child_lo, child_bound; };
template <int D> int launch_import(const yr::fd::Args<D>& a, unsigned long long n, void* stream);
template <int D> int launch_tensor(const yr::fd::Args<D>& a, int op, int cls, const unsigned int* list, unsigned long long n,
                                   unsigned long long need, int ext, void* stream);
template <int D> int launch_scalar(const yr::fd::Args<D>& a, const Segments& s, void* stream);
// yr_sorted_rec<D>[n] in (system, path) order; returns a device buffer (or null)
template <int D> void* sort_results(const void* results, unsigned long long n, void* stream);
}  // namespace yr_fd_backend

namespace yr_fd_host {
using namespace yr;
using namespace yr::f2;
using yr_f2_host::fail;
using yr_f2_host::fail_no_axis;
using yr_f2_host::DevBuf;
using yr_f2_host::release;

template <int D>
struct Frontier {
    bool active = false;
    fd::Args<D> a; Ctl h;
    unsigned long long n0 = 0, n_boxes = 0, n_results = 0, n_nodes = 0, rounds = 0, launches = 0;
    unsigned long long result_base0 = 0, node_base0 = 0, in_used = 0;
    const double* data_in = nullptr;
    bool own_input = false;
    int cur = 0, out_arena = 0, in_arena = -1;
    size_t cap_boxes = 0, cap_results = 0, cap_nodes = 0, per_list[2] = {0, 0};
    int time_kernels = 0, nt = 0;
    bool trace = false, phases = false;
};
template <int D> inline Frontier<D>& frontier() { static Frontier<D> f; return f; }
template <int D> inline yr_f2_host::Workspace& workspace() { static yr_f2_host::Workspace w; return w; }
template <int D> inline int trim_workspace(void* stream) {
    yr_f2_host::Workspace& w = workspace<D>();
    w.recs.release(stream); w.work.release(stream); w.arena[0].release(stream); w.arena[1].release(stream);
    w.lists[0].release(stream); w.lists[1].release(stream); w.tabs.release(stream); w.ctlbuf.release(stream);
    w.results.release(stream); w.nodes.release(stream);
    w.tabs_nt = 0;
    return 0;
}

// 1 when boxes whose largest tensor has `max_tensor` coefficients and largest extent `max_extent` fit the shared-memory
// budget of both tensor ops (conservative: the last axis may be padded by one element per row)
template <int D>
inline int fits(unsigned long long max_extent, unsigned long long max_tensor) {
    if (max_extent > 4096 || max_tensor > (1ull << 24)) return 0;
    const unsigned long long T = max_tensor + max_tensor / 2 + 2;
    const int n = (int)(max_extent < 4 ? 4 : max_extent);
    const unsigned long long P = (unsigned long long)panel_doubles(n);
    return (D + 1) * T + 2 * P <= kMaxNeed && T + D * P + (unsigned long long)(D * n) <= kMaxNeed;
}

template <int D> inline void carve(Frontier<D>& f, DevBuf& buf, size_t per_list) {
    unsigned int* base = static_cast<unsigned int*>(buf.p);
    for (int o = 0; o < kOps; ++o) for (int c = 0; c < kClasses; ++c) f.a.lists[o][c] = base + (size_t)(o * kClasses + c) * per_list;
}
template <int D> inline void bind_tables(Frontier<D>& f) {
    yr_f2_host::Workspace& ws = workspace<D>();
    f.a.recs = static_cast<BoxRec<D>*>(ws.recs.p); f.a.work = static_cast<fd::Work<D>*>(ws.work.p); f.a.cap_boxes = f.cap_boxes;
    f.a.results = static_cast<ResultRec<D>*>(ws.results.p); f.a.cap_results = f.cap_results;
    f.a.nodes = static_cast<NodeRec<D>*>(ws.nodes.p); f.a.cap_nodes = f.cap_nodes;
}

#define F2_TRY(expr) do { if (expr) { f.active = false; return -1; } } while (0)

template <int D>
inline int begin(const yr_frontier_desc* d, void* stream) {
    namespace be = yr_f2_backend; namespace bd = yr_fd_backend;
    Frontier<D>& f = frontier<D>();
    if (f.active) return fail("yr_frontier_begin: a frontier solve is already in progress (one per process)");
    if (d->dim != D) return fail("yr_frontier_solve: dimension mismatch");
    if (d->prm.exact) return fail("yr_frontier_solve: exact mode is served by yr_process_level");
    if (d->prm.trim_policy) return fail("yr_frontier_solve: the alternative trim policies are served by yr_process_level");
    const unsigned long long n0 = d->n_in;
    if (n0 > 0x3FFFFFFFull) return fail("yr_frontier_solve: too many level-0 boxes");
    const int nt = (int)(d->max_extent < 4 ? 4 : d->max_extent);
    if (nt > 4096) return fail("yr_frontier_solve: tensors too large for the shared-memory pipeline");
    f = Frontier<D>();
    f.active = true; f.n0 = n0; f.nt = nt; f.time_kernels = d->time_kernels;
    be::timer_reset(d->time_kernels);
    f.trace = getenv("YR_F2_TRACE") != nullptr;
    f.phases = getenv("YR_F2_PHASES") != nullptr;       // coarse host-clock phase summary (synchronises)
    if (f.phases) be::trace_sync_ms(stream);

    // Box tables, arenas, queues, matrix tables and the control block persist between solves (grow-only, released
    // by yr_frontier_trim); result / node records accumulate in persistent buffers sized by the (loose) per-round
    // bounds, the caller gets exactly sized copies at the end.
    yr_f2_host::Workspace& ws = workspace<D>();
    // level-independent subdivision matrices
    const size_t P = (size_t)panel_doubles(nt);
    F2_TRY(ws.tabs.reserve(4 * P * 8 + 4 * (size_t)nt * 4 + 64, 0, stream, 1.0));
    double* mats = static_cast<double*>(ws.tabs.p);
    int* rows = reinterpret_cast<int*>(mats + 4 * P);
    if (ws.tabs_nt != nt) { F2_TRY(be::launch_tables(mats, rows, nt, stream)); ws.tabs_nt = nt; }

    F2_TRY(ws.ctlbuf.reserve(sizeof(Ctl), 0, stream, 1.0));
    F2_TRY(be::dev_zero(ws.ctlbuf.p, sizeof(Ctl), stream));
    f.cap_boxes = (size_t)(n0 * 8 + 4096);
    F2_TRY(ws.recs.reserve(f.cap_boxes * sizeof(BoxRec<D>), 0, stream, 1.0));
    F2_TRY(ws.work.reserve(f.cap_boxes * sizeof(fd::Work<D>), 0, stream, 1.0));
    { const size_t have = ws.recs.cap / sizeof(BoxRec<D>) < ws.work.cap / sizeof(fd::Work<D>) ? ws.recs.cap / sizeof(BoxRec<D>) : ws.work.cap / sizeof(fd::Work<D>); if (have > f.cap_boxes) f.cap_boxes = have; }
    if (n0) F2_TRY(be::dev_copy(ws.recs.p, d->recs_in, n0 * sizeof(BoxRec<D>), stream));
    f.cap_results = (size_t)(n0 * 4 + 1024); f.cap_nodes = f.cap_results;
    F2_TRY(ws.results.reserve(f.cap_results * sizeof(ResultRec<D>), 0, stream, 1.0));
    F2_TRY(ws.nodes.reserve(f.cap_nodes * sizeof(NodeRec<D>), 0, stream, 1.0));
    if (ws.results.cap / sizeof(ResultRec<D>) > f.cap_results) f.cap_results = ws.results.cap / sizeof(ResultRec<D>);
    if (ws.nodes.cap / sizeof(NodeRec<D>) > f.cap_nodes) f.cap_nodes = ws.nodes.cap / sizeof(NodeRec<D>);

    fd::Args<D>& a = f.a;
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
        a.pol.warp_max[1] = env_ull("YR_F2_WARP_MAX_S", 2048ull);    // SUBDIVIDE: 64-thread CTA teams above (see launch_tensor_t)
        a.pol.cta128_max = env_ull("YR_F2_CTA128_MAX", 7168ull);
        for (int o = 0; o < 2; ++o) if (a.pol.warp_max[o] > kWarp32Max) a.pol.warp_max[o] = kWarp32Max;
        if (a.pol.cta128_max > 7168ull) a.pol.cta128_max = 7168ull;
    }
    memset(&f.h, 0, sizeof(f.h));
    // round -1: adopt the level-0 boxes (they arrive packed, from yr_init_boxes) and queue their first pass
    f.cur = 0;
    f.per_list[0] = (size_t)(n0 ? n0 : 1);
    F2_TRY(ws.lists[0].reserve((size_t)kOps * kClasses * f.per_list[0] * 4 + 64, 0, stream, 1.25));
    carve<D>(f, ws.lists[0], f.per_list[0]);
    bind_tables<D>(f);
    a.data_in = d->data_in;
    f.data_in = d->data_in; f.own_input = false; f.in_arena = -1; f.in_used = 0;
    f.n_boxes = n0; f.launches = 2;
    if (n0) {
        F2_TRY(bd::launch_import<D>(a, n0, stream));
        F2_TRY(be::to_host(&f.h, ws.ctlbuf.p, sizeof(Ctl), stream));
    }
    f.out_arena = 0;
    return 0;
}

// runs rounds until the queues are empty or `max_rounds` rounds have run (0: no limit)
template <int D>
inline int advance(unsigned long long max_rounds, void* stream) {
    namespace be = yr_f2_backend; namespace bd = yr_fd_backend;
    Frontier<D>& f = frontier<D>();
    if (!f.active) return fail("yr_frontier_advance: no frontier solve in progress");
    yr_f2_host::Workspace& ws = workspace<D>();
    fd::Args<D>& a = f.a; Ctl& h = f.h;
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
        const unsigned long long produced = n_res + (1ull << D) * n_sub;            // boxes the scalar step of this round may see
        for (int o = 0; o < kOps; ++o) for (int c = 0; c < kClasses; ++c) cur_lists[o][c] = a.lists[o][c];
        // capacities for everything this round can append
        if (f.n_boxes + (1ull << D) * n_sub > f.cap_boxes) {
            const size_t want = (size_t)((f.n_boxes + (1ull << D) * n_sub) * 2) + 4096;
            F2_TRY(ws.recs.reserve(want * sizeof(BoxRec<D>), f.n_boxes * sizeof(BoxRec<D>), stream, 1.0));
            F2_TRY(ws.work.reserve(want * sizeof(fd::Work<D>), f.n_boxes * sizeof(fd::Work<D>), stream, 1.0));
            f.cap_boxes = want;
        }
        if (f.n_results + produced > f.cap_results) {
            const size_t want = (size_t)((f.n_results + produced) * 2) + 1024;
            F2_TRY(ws.results.reserve(want * sizeof(ResultRec<D>), f.n_results * sizeof(ResultRec<D>), stream, 1.0));
            f.cap_results = want;
        }
        if (f.n_nodes + produced > f.cap_nodes) {
            const size_t want = (size_t)((f.n_nodes + produced) * 2) + 1024;
            F2_TRY(ws.nodes.reserve(want * sizeof(NodeRec<D>), f.n_nodes * sizeof(NodeRec<D>), stream, 1.0));
            f.cap_nodes = want;
        }
        F2_TRY(ws.arena[f.out_arena].reserve((size_t)(h.out_bound + 16) * 8, 0, stream, 1.25));
        f.per_list[f.cur ^ 1] = (size_t)produced;
        F2_TRY(ws.lists[f.cur ^ 1].reserve((size_t)kOps * kClasses * f.per_list[f.cur ^ 1] * 4 + 64, 0, stream, 1.25));
        carve<D>(f, ws.lists[f.cur ^ 1], f.per_list[f.cur ^ 1]);
        bind_tables<D>(f);
        a.results = static_cast<ResultRec<D>*>(ws.results.p) + f.n_results; a.cap_results = f.cap_results - f.n_results;
        a.nodes = static_cast<NodeRec<D>*>(ws.nodes.p) + f.n_nodes; a.cap_nodes = f.cap_nodes - f.n_nodes;
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
                F2_TRY(bd::launch_tensor<D>(a, o == 0 ? (int)kTRestrict : (int)kTSubdivide, c, cur_lists[o][c], h.n_list[o][c],
                                         h.need_max[o][c], (int)(o ? h.ext_max[c] : 0), stream));
                ++f.launches;
                if (f.trace) fprintf(stderr, " %s%d n=%llu need=%llu ext=%llu %.3f ms |", o ? "S" : "R", c, h.n_list[o][c], h.need_max[o][c], o ? h.ext_max[c] : 0ull, be::trace_sync_ms(stream));
            }
        be::round_join(stream);
        be::timer_mark(0, 0, stream);
        bd::Segments sg;
        for (int c = 0; c < kClasses; ++c) { sg.list[c] = cur_lists[0][c]; sg.n[c] = h.n_list[0][c]; }
        sg.child_lo = f.n_boxes; sg.child_bound = (1ull << D) * n_sub;
        be::timer_mark(1, 1, stream);
        F2_TRY(bd::launch_scalar<D>(a, sg, stream));
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

template <int D>
inline int finish(yr_frontier_out* out, void* stream) {
    namespace be = yr_f2_backend; namespace bd = yr_fd_backend;
    Frontier<D>& f = frontier<D>();
    memset(out, 0, sizeof(*out));
    if (!f.active) return fail("yr_frontier_finish: no frontier solve in progress");
    yr_f2_host::Workspace& ws = workspace<D>();
    f.active = false;
    const unsigned long long n_results = f.n_results, n_nodes = f.n_nodes;
    double ph2 = 0.0;
    out->order = n_results ? bd::sort_results<D>(ws.results.p, n_results, stream) : nullptr;
    if (f.phases) { ph2 = be::trace_sync_ms(stream); fprintf(stderr, "f2 phases: sort %.2f ms (%llu rounds, %llu launches)\n", ph2, f.rounds, f.launches); }
    if (n_results) {
        out->results = be::dev_alloc((size_t)n_results * sizeof(ResultRec<D>), stream);
        if (!out->results || be::dev_copy(out->results, ws.results.p, (size_t)n_results * sizeof(ResultRec<D>), stream)) { release(out, stream); yr_set_error_oom("device allocation failed (results)"); return -1; }
    }
    if (n_nodes) {
        out->nodes = be::dev_alloc((size_t)n_nodes * sizeof(NodeRec<D>), stream);
        if (!out->nodes || be::dev_copy(out->nodes, ws.nodes.p, (size_t)n_nodes * sizeof(NodeRec<D>), stream)) { release(out, stream); yr_set_error_oom("device allocation failed (nodes)"); return -1; }
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

template <int D>
inline int solve(const yr_frontier_desc* d, yr_frontier_out* out, void* stream) {
    memset(out, 0, sizeof(*out));
    if (d->n_in == 0) return 0;
    if (begin<D>(d, stream)) return -1;
    if (advance<D>(0, stream)) return -1;
    return finish<D>(out, stream);
}
#undef F2_TRY

}  // namespace yr_fd_host
