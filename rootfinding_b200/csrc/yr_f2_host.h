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
struct Segments { const unsigned int* list[yr::f2::kClasses]; unsigned long long n[yr::f2::kClasses];
                  unsigned long long child_lo, child_bound; };
int launch_scalar(const yr::f2::Args& a, const Segments& s, void* stream);
// optional CUDA-event timing of the tensor / scalar launches of a solve
void timer_reset(int enable);
void timer_mark(int which, int begin, void* stream);      // which: 0 tensor, 1 scalar
void timer_collect(double* tensor_ms, double* scalar_ms);
// order[0..n) = record indices sorted by (system, path_hi, path_lo); returns a device buffer (or null)
void* sort_results(const void* results, unsigned long long n, void* stream);
double trace_sync_ms(void* stream);      // YR_F2_TRACE: synchronise and return host milliseconds since the previous call
double host_ms();                        // host clock, milliseconds
}  // namespace yr_f2_backend

namespace yr_f2_host {
using namespace yr;
using namespace yr::f2;

inline int fail(const char* msg) { yr_set_error(msg); return -1; }

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    // grow to at least `bytes`, keeping the first `keep` bytes
    int reserve(size_t bytes, size_t keep, void* stream, double slack = 1.5) {
        if (bytes <= cap) return 0;
        size_t want = (size_t)(bytes * slack) + 256;
        void* q = yr_f2_backend::dev_alloc(want, stream);
        if (!q) return fail("device allocation failed (frontier buffers)");
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

inline int solve(const yr_frontier_desc* d, yr_frontier_out* out, void* stream) {
    namespace be = yr_f2_backend;
    memset(out, 0, sizeof(*out));
    if (d->dim != 2) return fail("yr_frontier_solve: dim must be 2");
    if (d->prm.exact) return fail("yr_frontier_solve: exact mode is served by yr_process_level");
    const unsigned long long n0 = d->n_in;
    if (n0 == 0) return 0;
    if (n0 > 0x3FFFFFFFull) return fail("yr_frontier_solve: too many level-0 boxes");
    const int nt = (int)(d->max_extent < 4 ? 4 : d->max_extent);
    if (!fits(d->max_extent)) return fail("yr_frontier_solve: tensors too large for the shared-memory pipeline");
    be::timer_reset(d->time_kernels);
    const bool trace = getenv("YR_F2_TRACE") != nullptr;
    const bool phases = getenv("YR_F2_PHASES") != nullptr;       // coarse host-clock phase summary (synchronises 3 times)
    double ph[4] = {0, 0, 0, 0};
    if (phases) be::trace_sync_ms(stream);

    // Box tables, arenas, queues, matrix tables and the control block persist between solves (grow-only, released
    // by yr_frontier_trim); result / node arrays belong to the caller of this solve.
    Workspace& ws = workspace();
    DevBuf &recs = ws.recs, &work = ws.work, &tabs = ws.tabs, &ctlbuf = ws.ctlbuf;
    DevBuf (&arena)[2] = ws.arena; DevBuf (&lists)[2] = ws.lists;
    // result / node records accumulate in persistent buffers sized by the (loose) per-round bounds; the caller gets
    // exactly sized copies at the end
    DevBuf &results = ws.results, &nodes = ws.nodes;
    auto cleanup = [&](bool) {};
#define F2_TRY(expr) do { if (expr) { cleanup(false); return -1; } } while (0)

    // level-independent subdivision matrices
    const size_t P = (size_t)panel_doubles(nt);
    F2_TRY(tabs.reserve(4 * P * 8 + 4 * (size_t)nt * 4 + 64, 0, stream, 1.0));
    double* mats = static_cast<double*>(tabs.p);
    int* rows = reinterpret_cast<int*>(mats + 4 * P);
    if (ws.tabs_nt != nt) { F2_TRY(be::launch_tables(mats, rows, nt, stream)); ws.tabs_nt = nt; }

    F2_TRY(ctlbuf.reserve(sizeof(Ctl), 0, stream, 1.0));
    F2_TRY(be::dev_zero(ctlbuf.p, sizeof(Ctl), stream));
    size_t cap_boxes = (size_t)(n0 * 8 + 4096);
    F2_TRY(recs.reserve(cap_boxes * sizeof(BoxRec<2>), 0, stream, 1.0));
    F2_TRY(work.reserve(cap_boxes * sizeof(Work), 0, stream, 1.0));
    { const size_t have = recs.cap / sizeof(BoxRec<2>) < work.cap / sizeof(Work) ? recs.cap / sizeof(BoxRec<2>) : work.cap / sizeof(Work); if (have > cap_boxes) cap_boxes = have; }
    F2_TRY(be::dev_copy(recs.p, d->recs_in, n0 * sizeof(BoxRec<2>), stream));
    size_t cap_results = (size_t)(n0 * 4 + 1024), cap_nodes = cap_results;
    F2_TRY(results.reserve(cap_results * sizeof(ResultRec<2>), 0, stream, 1.0));
    F2_TRY(nodes.reserve(cap_nodes * sizeof(NodeRec<2>), 0, stream, 1.0));
    if (results.cap / sizeof(ResultRec<2>) > cap_results) cap_results = results.cap / sizeof(ResultRec<2>);
    if (nodes.cap / sizeof(NodeRec<2>) > cap_nodes) cap_nodes = nodes.cap / sizeof(NodeRec<2>);

    Args a;
    memset(&a, 0, sizeof(a));
    a.ctl = static_cast<Ctl*>(ctlbuf.p);
    a.tab.nt = nt;
    for (int m = 0; m < 2; ++m) for (int h = 0; h < 2; ++h) { a.tab.mat[m][h] = mats + (size_t)(2 * m + h) * P; a.tab.rows_after[m][h] = rows + (size_t)(2 * m + h) * nt; }
    memcpy(&a.prm, &d->prm, sizeof(a.prm));
    a.result_base = d->result_base; a.node_base = d->node_base;
    {
        // team policy: warp teams up to these needs (doubles), one CTA per box above; overridable for experiments
        auto env_ull = [](const char* name, unsigned long long dflt) { const char* v = getenv(name); return v ? strtoull(v, nullptr, 10) : dflt; };
        a.pol.warp_max[0] = env_ull("YR_F2_WARP_MAX_R", kWarp32Max);
        a.pol.warp_max[1] = env_ull("YR_F2_WARP_MAX_S", kWarp32Max);
        a.pol.cta128_max = env_ull("YR_F2_CTA128_MAX", 7168ull);
        for (int o = 0; o < 2; ++o) if (a.pol.warp_max[o] > kWarp32Max) a.pol.warp_max[o] = kWarp32Max;
        if (a.pol.cta128_max > 7168ull) a.pol.cta128_max = 7168ull;
    }

    Ctl h;                               // host copy of the control block
    memset(&h, 0, sizeof(h));
    auto carve_lists = [&](DevBuf& buf, size_t per_list) -> int {
        if (buf.reserve((size_t)kOps * kClasses * per_list * 4 + 64, 0, stream, 1.25)) return -1;
        unsigned int* base = static_cast<unsigned int*>(buf.p);
        for (int o = 0; o < kOps; ++o) for (int c = 0; c < kClasses; ++c) a.lists[o][c] = base + (size_t)(o * kClasses + c) * per_list;
        return 0;
    };
    auto bind_tables = [&]() {
        a.recs = static_cast<BoxRec<2>*>(recs.p); a.work = static_cast<Work*>(work.p); a.cap_boxes = cap_boxes;
        a.results = static_cast<ResultRec<2>*>(results.p); a.cap_results = cap_results;
        a.nodes = static_cast<NodeRec<2>*>(nodes.p); a.cap_nodes = cap_nodes;
    };

    if (phases) ph[0] = be::trace_sync_ms(stream);
    // round -1: adopt the level-0 boxes (they arrive packed, from yr_init_boxes) and queue their first pass
    int cur = 0;
    F2_TRY(carve_lists(lists[cur], (size_t)n0));
    bind_tables();
    a.data_in = d->data_in;
    F2_TRY(be::launch_import(a, n0, stream));
    F2_TRY(be::to_host(&h, ctlbuf.p, sizeof(Ctl), stream));
    unsigned long long n_boxes = n0, n_results = 0, n_nodes = 0;
    const double* data_in = d->data_in;
    unsigned int* cur_lists[kOps][kClasses];
    unsigned long long rounds = 0, launches = 2;
    int out_arena = 0;
    for (;;) {
        unsigned long long n_res = 0, n_sub = 0;
        for (int c = 0; c < kClasses; ++c) { n_res += h.n_list[0][c]; n_sub += h.n_list[1][c]; }
        if (n_res + n_sub == 0) break;
        const double t_round0 = phases ? be::host_ms() : 0.0;
        if (++rounds > 1000000ull) { cleanup(false); return fail("yr_frontier_solve: round loop did not terminate"); }
        const unsigned long long produced = n_res + 4 * n_sub;            // boxes the scalar step of this round may see
        for (int o = 0; o < kOps; ++o) for (int c = 0; c < kClasses; ++c) cur_lists[o][c] = a.lists[o][c];
        // capacities for everything this round can append
        if (n_boxes + 4 * n_sub > cap_boxes) {
            const size_t want = (size_t)((n_boxes + 4 * n_sub) * 2) + 4096;
            F2_TRY(recs.reserve(want * sizeof(BoxRec<2>), n_boxes * sizeof(BoxRec<2>), stream, 1.0));
            F2_TRY(work.reserve(want * sizeof(Work), n_boxes * sizeof(Work), stream, 1.0));
            cap_boxes = want;
        }
        if (n_results + produced > cap_results) {
            const size_t want = (size_t)((n_results + produced) * 2) + 1024;
            F2_TRY(results.reserve(want * sizeof(ResultRec<2>), n_results * sizeof(ResultRec<2>), stream, 1.0));
            cap_results = want;
        }
        if (n_nodes + produced > cap_nodes) {
            const size_t want = (size_t)((n_nodes + produced) * 2) + 1024;
            F2_TRY(nodes.reserve(want * sizeof(NodeRec<2>), n_nodes * sizeof(NodeRec<2>), stream, 1.0));
            cap_nodes = want;
        }
        F2_TRY(arena[out_arena].reserve((size_t)(h.out_bound + 16) * 8, 0, stream, 1.25));
        F2_TRY(carve_lists(lists[cur ^ 1], (size_t)produced));
        bind_tables();
        a.results = static_cast<ResultRec<2>*>(results.p) + n_results; a.cap_results = cap_results - n_results;
        a.nodes = static_cast<NodeRec<2>*>(nodes.p) + n_nodes; a.cap_nodes = cap_nodes - n_nodes;
        a.result_base = d->result_base + n_results; a.node_base = d->node_base + n_nodes;
        a.coarse = (produced < 8192ull) ? 1 : 0;
        a.data_in = data_in; a.data_out = static_cast<double*>(arena[out_arena].p); a.cap_data_out = arena[out_arena].cap / 8;
        // per-round part of the control block: queues, bounds, arena cursor; slot counters restart at their totals
        F2_TRY(be::dev_zero(ctlbuf.p, offsetof(Ctl, n_boxes), stream));
        F2_TRY(be::dev_zero(&a.ctl->n_results, 2 * sizeof(unsigned long long), stream));
        if (trace) fprintf(stderr, "f2 round %3llu: host %.3f ms |", rounds, be::trace_sync_ms(stream));
        be::timer_mark(0, 1, stream);
        be::round_fork(stream);
        for (int o = kOps - 1; o >= 0; --o)
            for (int c = 0; c < kClasses; ++c) {
                if (!h.n_list[o][c]) continue;
                if (h.need_max[o][c] + (o ? 2ull * panel_doubles((int)(h.ext_max[c] < 4 ? 4 : h.ext_max[c])) : 0ull) > kMaxNeed) { cleanup(false); return fail("yr_frontier_solve: a box exceeds the shared-memory budget"); }
                F2_TRY(be::launch_tensor(a, o == 0 ? (int)kTRestrict : (int)kTSubdivide, c, cur_lists[o][c], h.n_list[o][c],
                                         h.need_max[o][c], (int)(o ? h.ext_max[c] : 0), stream));
                ++launches;
                if (trace) fprintf(stderr, " %s%d n=%llu need=%llu ext=%llu %.3f ms |", o ? "S" : "R", c, h.n_list[o][c], h.need_max[o][c], o ? h.ext_max[c] : 0ull, be::trace_sync_ms(stream));
            }
        be::round_join(stream);
        be::timer_mark(0, 0, stream);
        be::Segments sg;
        for (int c = 0; c < kClasses; ++c) { sg.list[c] = cur_lists[0][c]; sg.n[c] = h.n_list[0][c]; }
        sg.child_lo = n_boxes; sg.child_bound = 4 * n_sub;
        be::timer_mark(1, 1, stream);
        F2_TRY(be::launch_scalar(a, sg, stream));
        be::timer_mark(1, 0, stream);
        ++launches;
        if (trace) fprintf(stderr, " scalar %.3f ms\n", be::trace_sync_ms(stream));
        const double t_issue = phases ? be::host_ms() : 0.0;
        F2_TRY(be::to_host(&h, ctlbuf.p, sizeof(Ctl), stream));
        if (phases) { const double t_done = be::host_ms(); fprintf(stderr, " r%llu:%.2f+%.2f", rounds, t_issue - t_round0, t_done - t_issue); }
        if (h.overflow) { cleanup(false); return fail("yr_frontier_solve: internal buffer overflow"); }
        if (h.bad_mid) { cleanup(false); return fail("yr_frontier_solve: unexpected split point (nextTransformPoints)"); }
        n_boxes = h.n_boxes; n_results += h.n_results; n_nodes += h.n_nodes;
        data_in = a.data_out;
        out_arena ^= 1; cur ^= 1;
    }
    out->order = n_results ? be::sort_results(results.p, n_results, stream) : nullptr;
    if (phases) { ph[2] = be::trace_sync_ms(stream); fprintf(stderr, "f2 phases: setup %.2f ms, sort %.2f ms (%llu rounds, %llu launches)\n", ph[0], ph[2], rounds, launches); }
    if (n_results) {
        out->results = be::dev_alloc((size_t)n_results * sizeof(ResultRec<2>), stream);
        if (!out->results || be::dev_copy(out->results, results.p, (size_t)n_results * sizeof(ResultRec<2>), stream)) { release(out, stream); return fail("device allocation failed (results)"); }
    }
    if (n_nodes) {
        out->nodes = be::dev_alloc((size_t)n_nodes * sizeof(NodeRec<2>), stream);
        if (!out->nodes || be::dev_copy(out->nodes, nodes.p, (size_t)n_nodes * sizeof(NodeRec<2>), stream)) { release(out, stream); return fail("device allocation failed (nodes)"); }
    }
    out->n_results = n_results; out->n_nodes = n_nodes;
    out->rounds = rounds; out->launches = launches; out->slots = n_boxes;
    yr_counters& c = out->counters;
    c.n_results = n_results; c.n_nodes = n_nodes; c.boxes = h.boxes; c.zooms = h.zooms; c.subdivisions = h.subdivisions;
    c.discard_const = h.dconst; c.discard_quad = h.dquad; c.discard_zoom = h.dzoom; c.entry_coeffs = h.entry_coeffs;
    c.restrict_macs = h.restrict_macs; c.max_level = h.max_level;
    be::timer_collect(&out->tensor_ms, &out->scalar_ms);
    cleanup(true);
    return 0;
#undef F2_TRY
}

}  // namespace yr_f2_host
