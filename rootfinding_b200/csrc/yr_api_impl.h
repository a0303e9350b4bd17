// extern "C" entry points of include/yroots_b200.h, written against a handful of
// backend hooks (namespace yr_backend) that the translation unit including this
// file provides: yr_kernels.cu (CUDA, the product) or tests/emul/emul.cpp (pthread
// emulator of the same kernel source, test infrastructure only).
#pragma once
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include "../../include/yroots_b200.h"
#include "yr_box.h"
#include "yr_phases.h"
#include "yr_items.h"
#include "yr_approx.h"

// Hooks every backend defines after including this header.
namespace yr_backend {
const char* build_info();
int device_limits(int32_t* sm_count, int64_t* max_smem_per_cta, int64_t* smem_per_sm);
template <int D> int launch_init(const yr::InitArgs<D>& a, const yr_launch& l, void* stream);
// phase-split pipeline: one step over pa.list_in[0 .. pa.n_list)
template <int D> int launch_tensor_step(const yr::PhaseArgs<D>& pa, const yr_launch& l, void* stream);
template <int D> int launch_scalar_step(const yr::PhaseArgs<D>& pa, void* stream);
template <int D> int launch_subdivide_step(const yr::PhaseArgs<D>& pa, const yr_launch& l, void* stream);
int read_words(const void* dev, unsigned long long* out, int count, void* stream);   // device -> host, synchronises the stream
// launch shape of one step of the phase-split pipeline for boxes of at most the given size
template <int D> int plan_step(int subdivide, unsigned long long n_items, unsigned long long total, unsigned long long tensor,
                               unsigned long long extent, int exact, const yr_launch& base, yr_launch* out);
int zero_bytes(void* dev, size_t bytes, void* stream);
template <int D> int launch_bils(int64_t n, const double* lin, const double* c0, const double* sumabs, const double* errs,
                                 const int32_t* final_step, double* interval, int32_t* flags, void* stream);
template <int D> int launch_quad(int64_t n, const double* coeffs, const int64_t* off, const int32_t* shapes,
                                 const double* tol, int32_t nd_check, int32_t* out, void* stream);
int launch_dct1(const double* values, double* coeffs, int64_t outer, int64_t npts, int64_t inner, void* stream);
int launch_upper_axis(const double* coef, const double* P, double* out, int64_t outer, int64_t n, int64_t inner, void* stream);
int launch_decay(int mode, int dim, int axis, const yr::DecayTensor& t, const yr::DecayTensor* prev, const double* values, int64_t nvalues,
                 double prev_supnorm, double rel_tol, double abs_tol, double* scratch, yr::DecayOut* out, void* stream);
int launch_poly_eval(const double* coef, const double* x, double* out, int64_t outer, int64_t ncoef, int64_t npts,
                     int64_t inner, int32_t basis, void* stream);
int launch_abs_axis_mean(const double* t, double* out, double* supnorm, int64_t outer, int64_t n, int64_t inner, void* stream);
}  // namespace yr_backend

static_assert(sizeof(yr_params) == sizeof(yr::SolverParams), "yr_params / SolverParams mismatch");
static_assert(sizeof(yr_counters) == sizeof(yr::LevelCounters), "yr_counters / LevelCounters mismatch");

namespace yr_api {
static thread_local char g_err[512] = "";
static thread_local int g_err_code = 0;                // yr_last_error_code: 1 generic, 2 out of device memory
inline int fail(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); g_err_code = 1; return -1; }

inline yr::SolverParams to_params(const yr_params& p) {
    yr::SolverParams q;
    memcpy(&q, &p, sizeof(q));
    return q;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <int D>
struct WorkLayout {
    size_t work, list_a, list_b, list_sub, ctl, total;
    explicit WorkLayout(unsigned long long n) {
        work = 0;
        list_a = align_up(sizeof(yr::BoxWork<D>) * n, 16);
        list_b = list_a + align_up(4 * n, 16);
        list_sub = list_b + align_up(4 * n, 16);
        ctl = list_sub + align_up(4 * n, 16);
        total = ctl + sizeof(yr::RoundCtl);
    }
};

// One level with the phase-split kernels: rounds of (tensor step, scalar step) over the shrinking list of
// boxes that still need work, then the subdivision step.  One 8-byte read-back per round.
template <int D>
int run_level_phased(const yr::LevelArgs<D>& a, const yr_level_desc* d, void* stream) {
    const WorkLayout<D> wl(a.n_in);
    if (!d->work || d->work_bytes < wl.total) return fail("work buffer missing or too small (yr_level_work_bytes)");
    if (a.n_in == 0) return 0;
    if (a.n_in > 0xFFFFFFF0ull) return fail("level too large for 32-bit box lists");
    char* base = static_cast<char*>(d->work);
    yr::PhaseArgs<D> pa;
    pa.lv = a;
    pa.work = reinterpret_cast<yr::BoxWork<D>*>(base + wl.work);
    unsigned int* lists[2] = {reinterpret_cast<unsigned int*>(base + wl.list_a), reinterpret_cast<unsigned int*>(base + wl.list_b)};
    pa.list_sub = reinterpret_cast<unsigned int*>(base + wl.list_sub);
    pa.ctl = reinterpret_cast<yr::RoundCtl*>(base + wl.ctl);
    unsigned long long sizes[3] = {d->max_box_doubles, d->max_tensor_doubles, d->max_extent};
    yr_launch L;
    if (!d->resume_subdivide) {
        if (yr_backend::zero_bytes(pa.ctl, sizeof(yr::RoundCtl), stream)) return -1;
        pa.list_in = nullptr; pa.n_list = a.n_in; pa.first_round = 1;
        int cur = 0;
        for (int round = 0; pa.n_list > 0; ++round) {
            if (round > 100000) return fail("round loop did not terminate");
            pa.list_next = lists[cur];
            if (yr_backend::plan_step<D>(0, pa.n_list, sizes[0], sizes[1], sizes[2], a.prm.exact, d->launch, &L)) return -1;
            pa.lv.gws = L.gws; pa.lv.gws_stride = L.gws_stride; pa.lv.ws_doubles = L.ws_doubles;
            if (yr_backend::launch_tensor_step<D>(pa, L, stream)) return -1;
            if (yr_backend::launch_scalar_step<D>(pa, stream)) return -1;
            unsigned long long w[6];                        // cursor, n_next, n_sub, max_total, max_tensor, max_extent
            if (yr_backend::read_words(pa.ctl, w, 6, stream)) return -1;
            if (yr_backend::zero_bytes(pa.ctl, 2 * sizeof(unsigned long long), stream)) return -1;            // cursor, n_next
            if (yr_backend::zero_bytes(&pa.ctl->max_total, 3 * sizeof(unsigned long long), stream)) return -1;
            pa.list_in = lists[cur]; pa.n_list = w[1]; pa.first_round = 0;
            sizes[0] = w[3]; sizes[1] = w[4]; sizes[2] = w[5];
            cur ^= 1;
        }
    }
    unsigned long long w[9];
    if (yr_backend::read_words(pa.ctl, w, 9, stream)) return -1;
    if (yr_backend::zero_bytes(&pa.ctl->cursor, sizeof(unsigned long long), stream)) return -1;
    const unsigned long long n_sub = w[2];
    pa.list_in = pa.list_sub; pa.n_list = n_sub; pa.first_round = 0; pa.list_next = nullptr;
    if (n_sub == 0) return 0;
    if (yr_backend::plan_step<D>(1, n_sub, w[6], w[7], w[8], a.prm.exact, d->launch, &L)) return -1;
    pa.lv.gws = L.gws; pa.lv.gws_stride = L.gws_stride; pa.lv.ws_doubles = L.ws_doubles;
    return yr_backend::launch_subdivide_step<D>(pa, L, stream);
}

template <int D>
int level_impl(const yr_level_desc* d, void* stream) {
    yr::LevelArgs<D> a;
    a.recs_in = static_cast<const yr::BoxRec<D>*>(d->recs_in); a.data_in = d->data_in; a.n_in = d->n_in;
    a.recs_out = static_cast<yr::BoxRec<D>*>(d->recs_out); a.data_out = d->data_out;
    a.cap_recs_out = d->cap_recs_out; a.cap_data_out = d->cap_data_out;
    a.results = static_cast<yr::ResultRec<D>*>(d->results); a.cap_results = d->cap_results;
    a.nodes = static_cast<yr::NodeRec<D>*>(d->nodes); a.cap_nodes = d->cap_nodes;
    a.result_base = d->result_base; a.node_base = d->node_base;
    a.ctr = static_cast<yr::LevelCounters*>(d->counters);
    a.gws = d->launch.gws; a.gws_stride = d->launch.gws_stride; a.ws_doubles = d->launch.ws_doubles;
    a.prm = to_params(d->prm);
    if (d->pipeline != 1) return fail("yr_process_level: pipeline must be 1 (phase-split kernels)");
    return run_level_phased<D>(a, d, stream);
}

template <int D>
int init_impl(const yr_init_desc* d, void* stream) {
    yr::InitArgs<D> a;
    a.coeffs = d->coeffs; a.coeff_off = reinterpret_cast<const long long*>(d->coeff_off);
    a.shapes = d->shapes; a.errs = d->errs;
    a.job_system = d->job_system; a.job_restrict = d->job_restrict;
    a.job_alpha = d->job_alpha; a.job_beta = d->job_beta; a.job_iv = d->job_iv;
    a.job_level = d->job_level; a.job_next_mid = d->job_next_mid; a.job_parent = d->job_parent;
    a.njobs = d->njobs;
    a.recs_out = static_cast<yr::BoxRec<D>*>(d->recs_out); a.data_out = d->data_out;
    a.cap_recs_out = d->cap_recs_out; a.cap_data_out = d->cap_data_out;
    a.ctr = static_cast<yr::LevelCounters*>(d->counters);
    a.gws = d->launch.gws; a.gws_stride = d->launch.gws_stride; a.ws_doubles = d->launch.ws_doubles;
    a.prm = to_params(d->prm);
    return yr_backend::launch_init<D>(a, d->launch, stream);
}

template <int D>
int64_t layout_impl(int kind, int64_t* t, int cap) {
    int n = 0;
    auto put = [&](int id, size_t off, size_t size) {
        if (t && n < cap) { t[3 * n] = id; t[3 * n + 1] = (int64_t)off; t[3 * n + 2] = (int64_t)size; }
        ++n;
    };
    typedef yr::BoxRec<D> B; typedef yr::ResultRec<D> R; typedef yr::NodeRec<D> N; typedef yr::IntervalState<D> S;
    if (kind == 0) {
        const size_t so = offsetof(B, st);
        put(0, offsetof(B, data_off), sizeof(uint64_t)); put(1, offsetof(B, shape), sizeof(int32_t) * D * D);
        put(2, offsetof(B, err), sizeof(double) * D);
        put(3, so + offsetof(S, iv), sizeof(double) * D * 2); put(4, so + offsetof(S, A), sizeof(double) * D * 2);
        put(5, so + offsetof(S, B), sizeof(double) * D * 2); put(6, so + offsetof(S, next_mid), sizeof(double) * D);
        put(7, so + offsetof(S, pf_iv), sizeof(double) * D * 2); put(8, so + offsetof(S, pfA), sizeof(double) * D * 2);
        put(9, so + offsetof(S, pfB), sizeof(double) * D * 2); put(10, so + offsetof(S, flags), sizeof(uint32_t));
        put(11, offsetof(B, system), 4); put(12, offsetof(B, parent_node), 4); put(13, offsetof(B, collector), 4);
        put(14, offsetof(B, level), 4); put(15, offsetof(B, path_hi), 8); put(16, offsetof(B, path_lo), 8);
        return (int64_t)sizeof(B);
    } else if (kind == 1) {
        put(0, offsetof(R, root), sizeof(double) * D); put(1, offsetof(R, box), sizeof(double) * D * 2);
        put(2, offsetof(R, comb_iv), sizeof(double) * D * 2); put(3, offsetof(R, iv_lo), sizeof(double) * D);
        put(4, offsetof(R, cell_iv), sizeof(double) * D * 2); put(5, offsetof(R, next_mid), sizeof(double) * D);
        put(6, offsetof(R, system), 4); put(7, offsetof(R, parent_node), 4); put(8, offsetof(R, collector), 4);
        put(9, offsetof(R, level), 4); put(10, offsetof(R, flags), 4); put(11, offsetof(R, path_hi), 8);
        put(12, offsetof(R, path_lo), 8);
        return (int64_t)sizeof(R);
    } else if (kind == 2) {
        put(0, offsetof(N, iv), sizeof(double) * D * 2); put(1, offsetof(N, next_mid), sizeof(double) * D);
        put(2, offsetof(N, system), 4); put(3, offsetof(N, parent_node), 4); put(4, offsetof(N, level), 4);
        put(5, offsetof(N, nsplit), 4);
        return (int64_t)sizeof(N);
    }
    return -1;
}
}  // namespace yr_api

#define YR_DISPATCH_DIM(dim, CALL)                                                   \
    switch (dim) {                                                                   \
        case 1: return CALL(1); case 2: return CALL(2); case 3: return CALL(3);      \
        case 4: return CALL(4); case 5: return CALL(5); case 6: return CALL(6);      \
        default: return yr_api::fail("dimension must be 1..6");                      \
    }

extern "C" {

int yr_abi_version(void) { return YR_ABI_VERSION; }
const char* yr_last_error(void) { return yr_api::g_err; }
void yr_set_error(const char* msg) { yr_api::fail(msg); }          /* used by the other translation units of the library */
void yr_set_error_oom(const char* msg) { yr_api::fail(msg); yr_api::g_err_code = 2; }
void yr_set_error_no_axis(void) {
    yr_api::fail("a box has no axis left to subdivide along (point-sized box at level <= 5); the reference raises here: "
                 "IndexError in getInverseOrder, ChebyshevSubdivisionSolver.py:1010");
    yr_api::g_err_code = 3;
}
int yr_last_error_code(void) { return yr_api::g_err_code; }
const char* yr_build_info(void) { return yr_backend::build_info(); }

void yr_default_params(yr_params* p) {
    memset(p, 0, sizeof(*p));
    p->exact = 0; p->constant_check = 1; p->low_dim_quad_check = 1; p->all_dim_quad_check = 0;
    p->max_zoom_count = yr::kMaxZoomCount; p->use_fma = 0; p->trim = 1;
    p->trim_rel_tol = 1e-3; p->trim_abs_tol = 0.0;
}

int64_t yr_record_layout(int32_t dim, int32_t kind, int64_t* triples, int32_t cap) {
#define YR_CALL(D) yr_api::layout_impl<D>(kind, triples, cap)
    YR_DISPATCH_DIM(dim, YR_CALL)
#undef YR_CALL
}

uint64_t yr_workspace_doubles(int32_t dim, uint64_t total, uint64_t max_tensor, uint64_t max_extent, int32_t exact) {
    return yr::workspace_doubles(dim, total, max_tensor, max_extent, exact);
}

uint64_t yr_level_work_bytes(int32_t dim, uint64_t n) {
    switch (dim) {
        case 1: return yr_api::WorkLayout<1>(n).total; case 2: return yr_api::WorkLayout<2>(n).total;
        case 3: return yr_api::WorkLayout<3>(n).total; case 4: return yr_api::WorkLayout<4>(n).total;
        case 5: return yr_api::WorkLayout<5>(n).total; case 6: return yr_api::WorkLayout<6>(n).total;
        default: return 0;
    }
}

int64_t yr_static_smem_bytes(int32_t dim) {
    switch (dim) {
        case 1: return sizeof(yr::BoxShared<1>) + 512; case 2: return sizeof(yr::BoxShared<2>) + 512;
        case 3: return sizeof(yr::BoxShared<3>) + 512; case 4: return sizeof(yr::BoxShared<4>) + 512;
        case 5: return sizeof(yr::BoxShared<5>) + 512; case 6: return sizeof(yr::BoxShared<6>) + 512;
        default: return -1;
    }
}

int yr_device_limits(int32_t* sm_count, int64_t* max_smem_per_cta, int64_t* smem_per_sm) {
    return yr_backend::device_limits(sm_count, max_smem_per_cta, smem_per_sm);
}

int yr_process_level(const yr_level_desc* d, void* stream) {
    if (!d) return yr_api::fail("null descriptor");
    if (d->launch.threads <= 0 || d->launch.threads % 32 != 0 || d->launch.ctas <= 0) return yr_api::fail("bad launch configuration");
#define YR_CALL(D) yr_api::level_impl<D>(d, stream)
    YR_DISPATCH_DIM(d->dim, YR_CALL)
#undef YR_CALL
}

int yr_init_boxes(const yr_init_desc* d, void* stream) {
    if (!d) return yr_api::fail("null descriptor");
    if (d->launch.threads <= 0 || d->launch.threads % 32 != 0 || d->launch.ctas <= 0) return yr_api::fail("bad launch configuration");
#define YR_CALL(D) yr_api::init_impl<D>(d, stream)
    YR_DISPATCH_DIM(d->dim, YR_CALL)
#undef YR_CALL
}

int yr_bils_batched(int32_t dim, int64_t n, const double* lin, const double* c0, const double* sumabs,
                    const double* errs, const int32_t* final_step, double* interval, int32_t* flags, void* stream) {
#define YR_CALL(D) yr_backend::launch_bils<D>(n, lin, c0, sumabs, errs, final_step, interval, flags, stream)
    YR_DISPATCH_DIM(dim, YR_CALL)
#undef YR_CALL
}

int yr_quadcheck_batched(int32_t dim, int64_t n, const double* coeffs, const int64_t* off, const int32_t* shapes,
                         const double* tol, int32_t nd_check, int32_t* out, void* stream) {
#define YR_CALL(D) yr_backend::launch_quad<D>(n, coeffs, off, shapes, tol, nd_check, out, stream)
    YR_DISPATCH_DIM(dim, YR_CALL)
#undef YR_CALL
}

int yr_dct1_axis(const double* values, double* coeffs, int64_t outer, int64_t npts, int64_t inner, void* stream) {
    if (npts < 2) return yr_api::fail("DCT-I needs at least two points");
    return yr_backend::launch_dct1(values, coeffs, outer, npts, inner, stream);
}

int yr_poly_eval_axis(const double* coef, const double* x, double* out, int64_t outer, int64_t ncoef,
                      int64_t npts, int64_t inner, int32_t basis, void* stream) {
    return yr_backend::launch_poly_eval(coef, x, out, outer, ncoef, npts, inner, basis, stream);
}

int yr_abs_axis_mean(const double* t, double* out, double* supnorm, int64_t outer, int64_t n, int64_t inner, void* stream) {
    return yr_backend::launch_abs_axis_mean(t, out, supnorm, outer, n, inner, stream);
}

int yr_upper_axis_apply(const double* coef, const double* P, double* out, int64_t outer, int64_t n, int64_t inner, void* stream) {
    if (n < 1) return yr_api::fail("yr_upper_axis_apply: empty axis");
    return yr_backend::launch_upper_axis(coef, P, out, outer, n, inner, stream);
}

int yr_decay_probe(int32_t mode, int32_t dim, int32_t axis, const double* coeff, const int64_t* shape, const int64_t* keep,
                   const double* prev_coeff, const int64_t* prev_shape, const int64_t* prev_keep, const double* values, int64_t nvalues,
                   double prev_supnorm, double rel_tol, double abs_tol, double* scratch, yr_decay_out* out, void* stream) {
    if (dim < 1 || dim > yr::kMaxApproxDim || axis < 0 || axis >= dim) return yr_api::fail("yr_decay_probe: bad dimension / axis");
    if (mode != 0 && mode != 1) return yr_api::fail("yr_decay_probe: mode must be 0 or 1");
    if (mode == 1 && !prev_coeff) return yr_api::fail("yr_decay_probe: mode 1 needs the previous probe");
    yr::DecayTensor t, p;
    memset(&t, 0, sizeof(t)); memset(&p, 0, sizeof(p));
    t.c = coeff; p.c = prev_coeff;
    for (int d = 0; d < dim; ++d) {
        t.shape[d] = shape[d]; t.keep[d] = keep[d];
        if (keep[d] < 1 || keep[d] > shape[d]) return yr_api::fail("yr_decay_probe: keep must be in 1..shape");
        if (mode == 1) { p.shape[d] = prev_shape[d]; p.keep[d] = prev_keep[d]; }
    }
    static_assert(sizeof(yr::DecayOut) == sizeof(yr_decay_out), "yr_decay_out layout");
    return yr_backend::launch_decay(mode, dim, axis, t, mode == 1 ? &p : nullptr, values, nvalues, prev_supnorm, rel_tol, abs_tol, scratch,
                                    reinterpret_cast<yr::DecayOut*>(out), stream);
}

}  // extern "C"

