// TEST INFRASTRUCTURE ONLY -- SPMD emulator of the CUDA kernel source.
//
// Compiles rootfinding_b200/csrc/yr_box.h (the per-box pipeline the CUDA kernels run)
// for the CPU and executes each "CTA" as a team of pthreads with a spin barrier, warp
// and block reductions performed in the same order as on the device.  It exports the
// same C ABI as the product library so tests can drive the real host code against it
// on a machine without a GPU.  The product package never loads this library.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../rootfinding_b200/csrc/yr_api_impl.h"
#include "../../rootfinding_b200/csrc/yr_approx.h"
#include "../../rootfinding_b200/csrc/yr_f2_host.h"
#include "../../rootfinding_b200/csrc/yr_fd_host.h"

namespace {

struct SpinBarrier {
    std::atomic<int> count{0};
    std::atomic<int> sense{0};
    int n = 1;
    void wait() {
        const int s = sense.load(std::memory_order_acquire);
        if (count.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
            count.store(0, std::memory_order_relaxed);
            sense.store(1 - s, std::memory_order_release);
        } else {
            int spins = 0;
            while (sense.load(std::memory_order_acquire) == s) { if (++spins > 2000) { std::this_thread::yield(); spins = 0; } }
        }
    }
};

struct Team {
    int T, lanes;
    SpinBarrier bar;
    std::vector<double> scratch;
};

struct CtaShared {          // what the warp-teams of one emulated CTA share
    SpinBarrier bar;
    std::atomic<int> vote{1};
};

template <bool WARP>
struct HostCtxT {
    static constexpr bool kWarpPerBox = WARP;
    int tid, nthreads, lane, lanes, warp, nwarps;
    Team* team;
    CtaShared* cta = nullptr;
    int cta_tid = 0;
    void sync() { team->bar.wait(); }
    void cta_sync() { if (cta) cta->bar.wait(); else team->bar.wait(); }
    bool cta_all(bool pred) {
        if (!cta) return pred;
        cta->bar.wait();
        if (cta_tid == 0) cta->vote.store(1);
        cta->bar.wait();
        if (!pred) cta->vote.store(0);
        cta->bar.wait();
        const bool r = cta->vote.load() != 0;
        cta->bar.wait();
        return r;
    }
    double tree(int w) const {
        double a[64];
        for (int l = 0; l < lanes; ++l) a[l] = team->scratch[w * lanes + l];
        for (int off = lanes / 2; off >= 1; off /= 2) for (int l = 0; l < off; ++l) a[l] += a[l + off];
        return a[0];
    }
    double block_sum(double v) {
        team->scratch[tid] = v; sync();
        double tot = 0; for (int w = 0; w < nwarps; ++w) tot += tree(w);
        sync(); return tot;
    }
    double block_max(double v) {
        team->scratch[tid] = v; sync();
        double m = team->scratch[0]; for (int t = 1; t < nthreads; ++t) m = team->scratch[t] > m ? team->scratch[t] : m;
        sync(); return m;
    }
    double block_max_nan(double v) {
        team->scratch[tid] = v; sync();
        double m = team->scratch[0];
        for (int t = 1; t < nthreads; ++t) { const double o = team->scratch[t]; m = (m != m || o != o) ? NAN : (o > m ? o : m); }
        sync(); return m;
    }
    double warp_sum(double v) {
        team->scratch[tid] = v; sync();
        const double r = tree(warp);
        sync(); return r;
    }
    // warp shuffles (collective over the team, which is one warp when WARP)
    double shfl_up(double v, int d) {
        team->scratch[tid] = v; sync();
        const double r = (lane >= d) ? team->scratch[tid - d] : v;
        sync(); return r;
    }
    double shfl_down(double v, int d) {
        team->scratch[tid] = v; sync();
        const double r = (lane + d < lanes) ? team->scratch[tid + d] : v;
        sync(); return r;
    }
    int shfl_idx_int(int v, int src) {
        team->scratch[tid] = (double)v; sync();
        const int r = (int)team->scratch[warp * lanes + src];
        sync(); return r;
    }
    unsigned long long atomic_add(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
    void atomic_max(unsigned long long* p, unsigned long long v) {
        unsigned long long cur = __atomic_load_n(p, __ATOMIC_RELAXED);
        while (cur < v && !__atomic_compare_exchange_n(p, &cur, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    }
    void copy_in(double* dst, const double* src, long long n) { for (long long i = tid; i < n; i += nthreads) dst[i] = src[i]; }
    double cospi_ratio(long long m, long long N) const { return cos(M_PI * (double)m / (double)N); }
};
typedef HostCtxT<false> HostCtx;
typedef HostCtxT<true> HostWarpCtx;

int env_int(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; }

template <class Body>
void run_team(int T, int lanes, Body body) {
    Team team; team.T = T; team.lanes = lanes; team.bar.n = T; team.scratch.assign(T, 0.0);
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t]() {
            HostCtx cx; cx.tid = t; cx.nthreads = T; cx.lanes = lanes; cx.lane = t % lanes; cx.warp = t / lanes;
            cx.nwarps = T / lanes; cx.team = &team;
            body(cx);
        });
    for (auto& x : th) x.join();
}

}  // namespace

namespace yr_backend {

const char* build_info() { return "emulator (pthreads; tests only)"; }

int device_limits(int32_t* sm_count, int64_t* max_smem_per_cta, int64_t* smem_per_sm) {
    if (sm_count) *sm_count = 2;
    if (max_smem_per_cta) *max_smem_per_cta = 227 * 1024;
    if (smem_per_sm) *smem_per_sm = 228 * 1024;
    return 0;
}

struct HostAtomics {
    unsigned long long add(unsigned long long* p, unsigned long long v) const { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
};

template <int D, class Body>
void run_box_teams(const yr_launch& l, Body body) {
    // tensor-type steps: CTA-per-box -> one team of YR_EMUL_THREADS; warp-per-box -> YR_EMUL_WARPS independent one-warp teams
    const int lanes = env_int("YR_EMUL_LANES", 4);
    const int W = l.warp_per_box ? env_int("YR_EMUL_WARPS", 2) : 1;
    const int T = l.warp_per_box ? lanes : env_int("YR_EMUL_THREADS", 8);
    std::vector<Team> teams(W);
    std::vector<std::vector<double>> wss(W);
    std::vector<yr::BoxShared<D>*> shs(W);
    for (int w = 0; w < W; ++w) {
        teams[w].T = T; teams[w].lanes = lanes; teams[w].bar.n = T; teams[w].scratch.assign(T, 0.0);
        if (l.ws_in_smem || l.warp_per_box) wss[w].assign(l.ws_doubles + 8, 0.0);
        shs[w] = new yr::BoxShared<D>();
    }
    std::vector<std::thread> th;
    for (int w = 0; w < W; ++w)
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, w, t]() {
                double* ws = wss[w].empty() ? l.gws + (size_t)w * l.gws_stride : wss[w].data();
                if (l.warp_per_box) {
                    HostWarpCtx cx; cx.tid = t; cx.nthreads = T; cx.lanes = lanes; cx.lane = t; cx.warp = 0; cx.nwarps = 1; cx.team = &teams[w];
                    body(cx, *shs[w], ws);
                } else {
                    HostCtx cx; cx.tid = t; cx.nthreads = T; cx.lanes = lanes; cx.lane = t % lanes; cx.warp = t / lanes;
                    cx.nwarps = T / lanes; cx.team = &teams[w];
                    body(cx, *shs[w], ws);
                }
            });
    for (auto& x : th) x.join();
    for (auto p : shs) delete p;
}

template <int D>
int launch_tensor_step(const yr::PhaseArgs<D>& pa, const yr_launch& l, void*) {
    run_box_teams<D>(l, [&](auto& cx, yr::BoxShared<D>& sh, double* ws) { yr::tensor_step_worker<D>(cx, pa, sh, ws); });
    return 0;
}

template <int D>
int launch_subdivide_step(const yr::PhaseArgs<D>& pa, const yr_launch& l, void*) {
    run_box_teams<D>(l, [&](auto& cx, yr::BoxShared<D>& sh, double* ws) { yr::subdivide_step_worker<D>(cx, pa, sh, ws); });
    return 0;
}

template <int D>
int launch_scalar_step(const yr::PhaseArgs<D>& pa, void*) {
    yr::ScalarCounts c{};
    for (unsigned long long i = 0; i < pa.n_list; ++i)
        yr::scalar_step<D>(pa, pa.list_in ? pa.list_in[i] : (unsigned int)i, c, HostAtomics());
    yr::LevelCounters* k = pa.lv.ctr;
    k->boxes += c.boxes; k->zooms += c.zooms; k->discard_const += c.dconst; k->discard_quad += c.dquad;
    k->discard_zoom += c.dzoom; k->entry_coeffs += c.coeffs;
    if (c.maxlevel > k->max_level) k->max_level = c.maxlevel;
    yr::RoundCtl* r = pa.ctl;
    if (c.t_total > r->max_total) r->max_total = c.t_total;
    if (c.t_tensor > r->max_tensor) r->max_tensor = c.t_tensor;
    if (c.t_extent > r->max_extent) r->max_extent = c.t_extent;
    if (c.s_total > r->sub_max_total) r->sub_max_total = c.s_total;
    if (c.s_tensor > r->sub_max_tensor) r->sub_max_tensor = c.s_tensor;
    if (c.s_extent > r->sub_max_extent) r->sub_max_extent = c.s_extent;
    return 0;
}

int read_words(const void* dev, unsigned long long* out, int count, void*) { memcpy(out, dev, 8 * (size_t)count); return 0; }

template <int D>
int plan_step(int subdivide, unsigned long long n_items, unsigned long long total, unsigned long long tensor,
              unsigned long long extent, int exact, const yr_launch& base, yr_launch* out) {
    *out = base;                                         // box granularity / placement as the test asked for
    out->ws_doubles = yr::workspace_doubles(D, total, tensor, extent, exact, subdivide);
    if (!base.ws_in_smem && !base.warp_per_box && out->ws_doubles > base.gws_stride) return yr_api::fail("global workspace too small");
    (void)n_items;
    return 0;
}
int zero_bytes(void* dev, size_t bytes, void*) { memset(dev, 0, bytes); return 0; }

template <int D>
int launch_init(const yr::InitArgs<D>& a, const yr_launch& l, void*) {
    const int T = env_int("YR_EMUL_THREADS", 8), lanes = env_int("YR_EMUL_LANES", 4);
    std::vector<double> ws;
    double* wsp;
    if (l.ws_in_smem) { ws.assign(l.ws_doubles + 8, 0.0); wsp = ws.data(); } else wsp = l.gws;
    yr::BoxShared<D>* sh = new yr::BoxShared<D>();
    run_team(T, lanes, [&](HostCtx& cx) { yr::init_worker<D>(cx, a, *sh, wsp); });
    delete sh;
    return 0;
}

template <int D>
int launch_bils(int64_t n, const double* lin, const double* c0, const double* sumabs, const double* errs,
                const int32_t* final_step, double* interval, int32_t* flags, void*) {
    for (int64_t i = 0; i < n; ++i) yr::bils_item<D>(i, lin, c0, sumabs, errs, final_step, interval, flags);
    return 0;
}

template <int D>
int launch_quad(int64_t n, const double* coeffs, const int64_t* off, const int32_t* shapes, const double* tol,
                int32_t nd_check, int32_t* out, void*) {
    for (int64_t i = 0; i < n; ++i)
        yr::quad_item<D>(i, coeffs, reinterpret_cast<const long long*>(off), shapes, tol, nd_check, out);
    return 0;
}

int launch_dct1(const double* values, double* coeffs, int64_t outer, int64_t npts, int64_t inner, void*) {
    yr::Dct1Args a{values, coeffs, outer, npts, inner};
    std::vector<double> table(2 * npts);
    run_team(4, 4, [&](HostCtx& cx) { yr::dct1_worker(cx, a, table.data(), 0, 1); });
    return 0;
}

int launch_poly_eval(const double* coef, const double* x, double* out, int64_t outer, int64_t ncoef, int64_t npts,
                     int64_t inner, int32_t basis, void*) {
    yr::PolyEvalArgs a{coef, x, out, outer, ncoef, npts, inner, basis};
    run_team(4, 4, [&](HostCtx& cx) { yr::poly_eval_worker(cx, a, 0, 1); });
    return 0;
}

int launch_upper_axis(const double* coef, const double* P, double* out, int64_t outer, int64_t n, int64_t inner, void*) {
    yr::UpperAxisArgs a{coef, P, out, outer, n, inner};
    run_team(4, 4, [&](HostCtx& cx) { yr::upper_axis_worker(cx, a, 0, 1); });
    return 0;
}

int launch_decay(int mode, int dim, int axis, const yr::DecayTensor& t, const yr::DecayTensor* prev, const double* values, int64_t nvalues,
                 double prev_supnorm, double rel_tol, double abs_tol, double* scratch, yr::DecayOut* out, void*) {
    const long long n = t.keep[axis];
    double* chunk = scratch; double* sup_part = scratch + n; double* diff_part = sup_part + 64;
    run_team(4, 4, [&](HostCtx& cx) { for (long long j = 0; j < n; ++j) yr::subbox_axis_mean_worker(cx, t, dim, axis, chunk, j); });
    run_team(4, 4, [&](HostCtx& cx) { yr::absmax_worker(cx, values, nvalues, sup_part, 0, 1); });
    if (mode == 1) run_team(4, 4, [&](HostCtx& cx) { yr::maxdiff_worker(cx, t, *prev, dim, diff_part, 0, 1); });
    yr::decay_decide(mode, chunk, n, sup_part, 1, diff_part, mode == 1 ? 1 : 0, prev_supnorm, rel_tol, abs_tol, out);
    return 0;
}

int launch_abs_axis_mean(const double* t, double* out, double* supnorm, int64_t outer, int64_t n, int64_t inner, void*) {
    yr::AbsMeanArgs a{t, out, supnorm, outer, n, inner};
    run_team(4, 4, [&](HostCtx& cx) { for (int64_t j = 0; j < n; ++j) yr::abs_mean_worker(cx, a, j); });
    return 0;
}

}  // namespace yr_backend

// ---- 2-D frontier pipeline (yr_f2.h) on pthreads ---------------------------------------------------
namespace {

struct F2Team {
    // the whole emulated team acts as one "warp" of `lanes` lanes
    static constexpr bool kShuffle = true;
    int tid, nthreads, lane, lanes, warp, warps;
    Team* team;
    void sync() { team->bar.wait(); }
    void wsync() { sync(); }
    double sum(double v) {
        team->scratch[tid] = v; sync();
        double t = 0; for (int i = 0; i < nthreads; ++i) t += team->scratch[i];
        sync(); return t;
    }
    double sum_sync(double v) { return sum(v); }
    void sum2_sync(double a, double b, double& ra, double& rb) { ra = sum(a); rb = sum(b); }
    double wsum(double v) { return sum(v); }
    bool wany(bool p) { return sum(p ? 1.0 : 0.0) > 0.0; }
    unsigned wballot(bool p) {
        team->scratch[tid] = p ? 1.0 : 0.0; sync();
        unsigned r = 0; for (int i = 0; i < nthreads && i < 32; ++i) if (team->scratch[i] != 0.0) r |= 1u << i;
        sync(); return r;
    }
    double wsum_w(double v, int w) {
        team->scratch[tid] = v; sync();
        double t = 0; const int lo = (tid / w) * w; for (int i = lo; i < lo + w && i < nthreads; ++i) t += team->scratch[i];
        sync(); return t;
    }
    double wshfl_idx_d(double v, int src) {
        team->scratch[tid] = v; sync();
        const double r = team->scratch[src];
        sync(); return r;
    }
    double wshfl_up(double v, int d, int w) {
        team->scratch[tid] = v; sync();
        const double r = ((tid % w) >= d) ? team->scratch[tid - d] : v;
        sync(); return r;
    }
    double wshfl_down(double v, int d, int w) {
        team->scratch[tid] = v; sync();
        const double r = ((tid % w) + d < w && tid + d < nthreads) ? team->scratch[tid + d] : v;
        sync(); return r;
    }
    int wshfl_idx(int v, int src, int w) {
        team->scratch[tid] = (double)v; sync();
        const int r = (int)team->scratch[(tid / w) * w + src];
        sync(); return r;
    }
    unsigned long long bcast(unsigned long long v) {
        if (tid == 0) memcpy(&team->scratch[0], &v, 8);
        sync();
        unsigned long long r; memcpy(&r, &team->scratch[0], 8);
        sync(); return r;
    }
    unsigned long long atomic_add(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
    void copy_async(double* dst, const double* src, int n_even) { for (int e = tid; e < n_even; e += nthreads) dst[e] = src[e]; }
    void copy_wait() {}
    void store_async(double* dst, const double* src, int n_even) { sync(); for (int e = tid; e < n_even; e += nthreads) dst[e] = src[e]; }
    void store_wait() {}
};

struct F2At {
    unsigned long long add(unsigned long long* p, unsigned long long v) const { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
    void max(unsigned long long* p, unsigned long long v) const { if (*p < v) *p = v; }
};

template <class Body>
void f2_run_team(int T, Body body) {
    Team team; team.T = T; team.lanes = T; team.bar.n = T; team.scratch.assign(T + 2, 0.0);
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t]() { F2Team tm; tm.tid = tm.lane = t; tm.nthreads = tm.lanes = T; tm.warp = 0; tm.warps = 1; tm.team = &team; body(tm); });
    for (auto& x : th) x.join();
}

void f2_flush(const yr::f2::Args& a, const yr::f2::Decision& dc) {
    yr::f2::Ctl* c = a.ctl;
    c->boxes += dc.boxes; c->zooms += dc.zooms; c->dconst += dc.dconst; c->dquad += dc.dquad; c->dzoom += dc.dzoom;
    c->entry_coeffs += dc.coeffs; c->subdivisions += dc.subdiv;
    if (dc.maxlevel > c->max_level) c->max_level = dc.maxlevel;
}

}  // namespace

namespace yr_f2_backend {
void* dev_alloc(size_t bytes, void*) { return malloc(bytes ? bytes : 16); }
void dev_free(void* p, void*) { free(p); }
int dev_copy(void* dst, const void* src, size_t bytes, void*) { memcpy(dst, src, bytes); return 0; }
int dev_zero(void* p, size_t bytes, void*) { memset(p, 0, bytes); return 0; }
int to_host(void* dst, const void* src, size_t bytes, void*) { memcpy(dst, src, bytes); return 0; }

int to_device(void* dst, const void* src, size_t bytes, void*) { memcpy(dst, src, bytes); return 0; }

int launch_export(const yr::f2::Args& a, const yr_f2_host::XTake& take, unsigned long long n, unsigned char* xbuf, const double* data_in, void*) {
    yr_f2_host::XHeader* hdr = reinterpret_cast<yr_f2_host::XHeader*>(xbuf);
    for (unsigned long long j = 0; j < n; ++j) {
        const unsigned int b = yr_f2_host::export_box_of(a, take, j);
        int sz0;
        const int tot = yr_f2_host::export_doubles(a, b, sz0);
        const unsigned long long at = hdr->data_doubles; hdr->data_doubles += (unsigned long long)tot;
        yr_f2_host::export_one(a, b, j, n, xbuf, data_in, 0, 1, at);
        yr_f2_host::export_fix_offsets(j, n, xbuf, at, sz0);
    }
    return 0;
}

int launch_import_x(const yr::f2::Args& a, unsigned long long n, const unsigned char* xbuf, unsigned long long slot_base, unsigned long long data_base, void*) {
    for (unsigned long long i = 0; i < n; ++i) {
        yr::f2::Decision dc; memset(&dc, 0, sizeof(dc));
        yr_f2_host::import_one(a, i, n, xbuf, (unsigned int)(slot_base + i), data_base, dc, F2At());
        yr::f2::commit_decision(a, (unsigned int)(slot_base + i), dc, F2At());
    }
    return 0;
}

int launch_tables(double* mats, int* rows, int nt, void*) {
    const int P = yr::f2::panel_doubles(nt);
    for (int m = 0; m < 4; ++m) {
        const double mid = (m >> 1) ? yr::kFirstSplit : 0.0;
        const double alpha = (mid + 1) / 2, beta = (mid - 1) / 2;
        const bool upper = m & 1;
        f2_run_team(4, [&](F2Team& tm) { yr::f2::build_panels(tm, mats + (size_t)m * P, rows + (size_t)m * nt, nt, upper ? -beta : alpha, upper ? alpha : beta); });
    }
    return 0;
}

int launch_import(const yr::f2::Args& a, unsigned long long n, void*) {
    for (unsigned long long i = 0; i < n; ++i) {
        yr::f2::Decision dc; memset(&dc, 0, sizeof(dc));
        yr::f2::import_box(a, (unsigned int)i, dc, F2At());
        yr::f2::commit_decision(a, (unsigned int)i, dc, F2At());
    }
    return 0;
}

int launch_tensor(const yr::f2::Args& a, int op, int cls, const unsigned int* list, unsigned long long n,
                  unsigned long long need, int ext, void*) {
    const int T = cls < yr::f2::kWarpClasses ? env_int("YR_EMUL_LANES", 4) : env_int("YR_EMUL_THREADS", 8);
    // the m = 0 matrices re-laid out for `ext`, as the CUDA kernels stage them in shared memory
    std::vector<double> smat;
    int snt = 0;
    if (op == yr::f2::kTSubdivide) {
        snt = ext < 4 ? 4 : ext;
        const int P = yr::f2::panel_doubles(snt), B = (snt + 3) >> 2;
        smat.assign(2 * (size_t)P, 0.0);
        for (int h = 0; h < 2; ++h)
            for (int b = 0; b < B; ++b)
                memcpy(&smat[(size_t)h * P + yr::f2::panel_off(b, snt)], a.tab.mat[0][h] + yr::f2::panel_off(b, a.tab.nt), (size_t)(snt - 4 * b) * 4 * 8);
    }
    std::vector<double> slice((size_t)need + 8, 0.0);
    const bool fma = a.prm.use_fma != 0;
    unsigned long long macs_total = 0;
    f2_run_team(T, [&](F2Team& tm) {
        unsigned long long macs = 0;
        for (unsigned long long i = 0; i < n; ++i) {
            const unsigned int b = list[i];
            const double* lo = smat.data(); const double* hi = smat.data() + yr::f2::panel_doubles(snt);
            if (cls < yr::f2::kWarpClasses) {                       // warp teams: in-place fiber-per-lane passes
                if (op == yr::f2::kTRestrict) { if (fma) yr::f2::restrict_box_w<true>(tm, a, b, slice.data(), macs); else yr::f2::restrict_box_w<false>(tm, a, b, slice.data(), macs); }
                else if (fma) yr::f2::subdivide_box_w<true>(tm, a, b, slice.data(), lo, hi, snt, macs); else yr::f2::subdivide_box_w<false>(tm, a, b, slice.data(), lo, hi, snt, macs);
            } else {
                if (op == yr::f2::kTRestrict) { if (fma) yr::f2::restrict_box<true>(tm, a, b, slice.data(), macs); else yr::f2::restrict_box<false>(tm, a, b, slice.data(), macs); }
                else if (fma) yr::f2::subdivide_box<true>(tm, a, b, slice.data(), lo, hi, snt, macs); else yr::f2::subdivide_box<false>(tm, a, b, slice.data(), lo, hi, snt, macs);
            }
            tm.sync();
        }
        if (tm.tid == 0) macs_total = macs;
    });
    a.ctl->restrict_macs += macs_total;
    return 0;
}

int launch_scalar(const yr::f2::Args& a, const Segments& s, void*) {
    auto one = [&](unsigned int b) {
        yr::f2::Decision dc; memset(&dc, 0, sizeof(dc));
        yr::f2::scalar_step(a, b, dc, F2At());
        yr::f2::commit_decision(a, b, dc, F2At());
        f2_flush(a, dc);
    };
    for (int c = 0; c < yr::f2::kClasses; ++c) for (unsigned long long i = 0; i < s.n[c]; ++i) one(s.list[c][i]);
    const unsigned long long hi = a.ctl->n_boxes;
    for (unsigned long long b = s.child_lo; b < hi; ++b) one((unsigned int)b);
    return 0;
}

int remap_links(void* results, unsigned long long nr, void* nodes, unsigned long long nn, const RankBases& rb, void*) {
    yr::ResultRec<2>* res = static_cast<yr::ResultRec<2>*>(results);
    yr::NodeRec<2>* nd = static_cast<yr::NodeRec<2>*>(nodes);
    const int mask = (1 << rb.shift) - 1;
    for (unsigned long long i = 0; i < nr; ++i) {
        const int p = res[i].parent_node, c = res[i].collector;
        if (p >= 0) res[i].parent_node = (int)(rb.node[p >> rb.shift] + (unsigned long long)(p & mask));
        if (c >= 0) res[i].collector = (int)(rb.res[c >> rb.shift] + (unsigned long long)(c & mask));
    }
    for (unsigned long long i = 0; i < nn; ++i) {
        const int p = nd[i].parent_node;
        if (p >= 0) nd[i].parent_node = (int)(rb.node[p >> rb.shift] + (unsigned long long)(p & mask));
    }
    return 0;
}
void* sort_results(const void* results, unsigned long long n, void*) {
    const yr::ResultRec<2>* res = static_cast<const yr::ResultRec<2>*>(results);
    std::vector<unsigned int> v((size_t)n);
    for (size_t i = 0; i < v.size(); ++i) v[i] = (unsigned int)i;
    std::stable_sort(v.begin(), v.end(), [&](unsigned int a, unsigned int b) {
        if (res[a].system != res[b].system) return res[a].system < res[b].system;
        if (res[a].path_hi != res[b].path_hi) return res[a].path_hi < res[b].path_hi;
        return res[a].path_lo < res[b].path_lo;
    });
    yr_sorted_rec* order = static_cast<yr_sorted_rec*>(malloc((size_t)n * sizeof(yr_sorted_rec) + 16));
    for (size_t i = 0; i < v.size(); ++i) {
        order[i].index = v[i]; order[i].system = res[v[i]].system; order[i].flags = res[v[i]].flags; order[i].collector = res[v[i]].collector;
        order[i].root[0] = res[v[i]].root[0]; order[i].root[1] = res[v[i]].root[1];
    }
    return order;
}
double trace_sync_ms(void*) { return 0.0; }
double host_ms() { return 0.0; }
void round_fork(void*) {}
void round_join(void*) {}
void* fan_stream(void* s) { return s; }
void timer_reset(int) {}
void timer_mark(int, int, void*) {}
void timer_collect(double* t, double* s) { *t = 0; *s = 0; }
}  // namespace yr_f2_backend

// ---- D-dimensional frontier pipeline (yr_fd.h) on pthreads ----------------------------------------
namespace yr_fd_backend {
void* fan_unused = nullptr;

template <int D>
int launch_import(const yr::fd::Args<D>& a, unsigned long long n, void*) {
    for (unsigned long long i = 0; i < n; ++i) {
        yr::f2::Decision dc; memset(&dc, 0, sizeof(dc));
        yr::fd::import_box<D>(a, (unsigned int)i, dc, F2At());
        yr::fd::commit_decision<D>(a, (unsigned int)i, dc, F2At());
    }
    return 0;
}

template <int D>
int launch_tensor(const yr::fd::Args<D>& a, int op, int cls, const unsigned int* list, unsigned long long n,
                  unsigned long long need, int ext, void*) {
    const int T = cls < yr::f2::kWarpClasses ? env_int("YR_EMUL_LANES", 4) : env_int("YR_EMUL_THREADS", 8);
    std::vector<double> smat;
    int snt = 0;
    if (op == yr::f2::kTSubdivide) {
        snt = ext < 4 ? 4 : ext;
        const int P = yr::f2::panel_doubles(snt), B = (snt + 3) >> 2;
        smat.assign(2 * (size_t)P, 0.0);
        for (int h = 0; h < 2; ++h)
            for (int b = 0; b < B; ++b)
                memcpy(&smat[(size_t)h * P + yr::f2::panel_off(b, snt)], a.tab.mat[0][h] + yr::f2::panel_off(b, a.tab.nt), (size_t)(snt - 4 * b) * 4 * 8);
    }
    std::vector<double> slice((size_t)need + 8, 0.0);
    const bool fma = a.prm.use_fma != 0;
    unsigned long long macs_total = 0;
    f2_run_team(T, [&](F2Team& tm) {
        unsigned long long macs = 0;
        for (unsigned long long i = 0; i < n; ++i) {
            const unsigned int b = list[i];
            const double* lo = smat.data(); const double* hi = smat.data() + yr::f2::panel_doubles(snt);
            if (op == yr::f2::kTRestrict) { if (fma) yr::fd::restrict_box<D, true>(tm, a, b, slice.data(), macs); else yr::fd::restrict_box<D, false>(tm, a, b, slice.data(), macs); }
            else { yr::fd::WalkState<D> walk; if (fma) yr::fd::subdivide_box<D, true>(tm, a, b, slice.data(), lo, hi, snt, macs, &walk); else yr::fd::subdivide_box<D, false>(tm, a, b, slice.data(), lo, hi, snt, macs, &walk); }
            tm.sync();
        }
        if (tm.tid == 0) macs_total = macs;
    });
    a.ctl->restrict_macs += macs_total;
    return 0;
}

template <int D>
int launch_scalar(const yr::fd::Args<D>& a, const Segments& s, void*) {
    auto one = [&](unsigned int b) {
        yr::f2::Decision dc; memset(&dc, 0, sizeof(dc));
        yr::fd::scalar_step<D>(a, b, dc, F2At());
        yr::fd::commit_decision<D>(a, b, dc, F2At());
        yr::f2::Ctl* c = a.ctl;
        c->boxes += dc.boxes; c->zooms += dc.zooms; c->dconst += dc.dconst; c->dquad += dc.dquad; c->dzoom += dc.dzoom;
        c->entry_coeffs += dc.coeffs; c->subdivisions += dc.subdiv;
        if (dc.maxlevel > c->max_level) c->max_level = dc.maxlevel;
    };
    for (int c = 0; c < yr::f2::kClasses; ++c) for (unsigned long long i = 0; i < s.n[c]; ++i) one(s.list[c][i]);
    const unsigned long long hi = a.ctl->n_boxes;
    for (unsigned long long b = s.child_lo; b < hi; ++b) one((unsigned int)b);
    return 0;
}

template <int D> struct SortedRec { uint32_t index; int32_t system; uint32_t flags; int32_t collector; double root[D]; };

template <int D>
void* sort_results(const void* results, unsigned long long n, void*) {
    const yr::ResultRec<D>* res = static_cast<const yr::ResultRec<D>*>(results);
    std::vector<unsigned int> v((size_t)n);
    for (size_t i = 0; i < v.size(); ++i) v[i] = (unsigned int)i;
    std::stable_sort(v.begin(), v.end(), [&](unsigned int a, unsigned int b) {
        if (res[a].system != res[b].system) return res[a].system < res[b].system;
        if (res[a].path_hi != res[b].path_hi) return res[a].path_hi < res[b].path_hi;
        return res[a].path_lo < res[b].path_lo;
    });
    SortedRec<D>* order = static_cast<SortedRec<D>*>(malloc((size_t)n * sizeof(SortedRec<D>) + 16));
    for (size_t i = 0; i < v.size(); ++i) {
        order[i].index = v[i]; order[i].system = res[v[i]].system; order[i].flags = res[v[i]].flags; order[i].collector = res[v[i]].collector;
        for (int d = 0; d < D; ++d) order[i].root[d] = res[v[i]].root[d];
    }
    return order;
}
}  // namespace yr_fd_backend

extern "C" {
int yr_frontier_solve(const yr_frontier_desc* d, yr_frontier_out* out, void* stream) {
    if (!d || !out) return yr_api::fail("null descriptor");
    if (d->dim == 3) return yr_fd_host::solve<3>(d, out, stream);
    if (d->dim == 4) return yr_fd_host::solve<4>(d, out, stream);
    return yr_f2_host::solve(d, out, stream);
}
int yr_frontier_fits_dim(int32_t dim, uint64_t max_extent, uint64_t max_tensor) {
    if (dim == 2) return yr_f2_host::fits(max_extent);
    if (dim == 3) return yr_fd_host::fits<3>(max_extent, max_tensor);
    if (dim == 4) return yr_fd_host::fits<4>(max_extent, max_tensor);
    return 0;
}
int yr_frontier_release(yr_frontier_out* out, void* stream) { return yr_f2_host::release(out, stream); }
int yr_gather_by_system(const void* table, uint64_t n, uint32_t rec_bytes, uint32_t system_off, const uint8_t* sys_flags, uint64_t nsys,
                        uint32_t* out_index, void* out_recs, uint64_t cap, uint64_t* out_n, void*) {
    const unsigned char* t = static_cast<const unsigned char*>(table);
    for (uint64_t i = 0; i < n; ++i) {
        int sys; memcpy(&sys, t + i * rec_bytes + system_off, 4);
        if (sys < 0 || (uint64_t)sys >= nsys || !sys_flags[sys]) continue;
        const uint64_t pos = (*out_n)++;
        if (pos >= cap) continue;
        out_index[pos] = (uint32_t)i;
        memcpy(static_cast<unsigned char*>(out_recs) + pos * rec_bytes, t + i * rec_bytes, rec_bytes);
    }
    return 0;
}
#include "../../rootfinding_b200/csrc/yr_f2_abi.inc"
int yr_frontier_fits(uint64_t max_extent) { return yr_f2_host::fits(max_extent); }
int yr_frontier_trim(void* stream) { yr_fd_host::trim_workspace<3>(stream); yr_fd_host::trim_workspace<4>(stream); return yr_f2_host::trim_workspace(stream); }
int yr_fetch(void* host_dst, const void* dev_src, uint64_t bytes, void*) { memcpy(host_dst, dev_src, bytes); return 0; }
}
