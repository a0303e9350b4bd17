// Plain-old-data records shared by the device kernels, the C-ABI and the
// Python host (which mirrors them as numpy structured dtypes; the layouts are
// exported at run time through yr_record_layout so nothing is hard-coded twice).
#pragma once
#include <stdint.h>
#include "yr_math.h"

namespace yr {

constexpr int kMaxDim = 6;

// Options that define the solver's behaviour.  Defaults == the reference's hard-coded
// constants ("Config / flags").
struct SolverParams {
    int32_t exact;                 // SolverOptions.exact
    int32_t constant_check;        // SolverOptions.constant_check
    int32_t low_dim_quad_check;    // SolverOptions.low_dim_quadratic_check
    int32_t all_dim_quad_check;    // SolverOptions.all_dim_quadratic_check
    int32_t max_zoom_count;        // 25
    int32_t use_fma;               // 0: separately rounded mul/add (bit-exact restriction); 1: fused
    int32_t trim;                  // 1: run trimMs (reference behaviour)
    int32_t trim_policy;           // 0: trimMs :1131-1171; 1-3: trimMsOptimized1/2/3 of the reference's experiment file (level pipeline)
    double  trim_rel_tol;          // 1e-3
    double  trim_abs_tol;          // 0
};

// One frontier box: D coefficient tensors (contiguous, row-major, back to back in the
// level's data pool) plus everything TrackedInterval carries.
template <int D>
struct BoxRec {
    uint64_t data_off;             // first double of tensor 0 in the pool
    int32_t  shape[D][D];          // shape[p][axis]
    double   err[D];               // approximation + accumulated transformation error per polynomial
    IntervalState<D> st;
    int32_t  system;               // which system of the batch
    int32_t  parent_node;          // NodeRec index of the subdividing parent (-1: top)
    int32_t  collector;            // ResultRec index of the enclosing final-step subdivision (-1: none)
    int32_t  level;                // solverOptions.level of the parent invocation
    int32_t  path_bits;            // bits of path_hi:path_lo in use
    int32_t  pad0;
    uint64_t path_hi, path_lo;     // child slots from the top, left aligned (MSB first): integer order == depth-first order
};

enum : uint32_t {
    kResFinalStep     = 1u << 0,   // box was in its final step when emitted
    kResExtraRoot     = 1u << 1,   // possibleExtraRoot: final-step subdivision found nothing (set by host pass)
    kResCollector     = 1u << 2,   // this record is a final-step subdivision parent; its roots come from its subtree
    kResTouchesCell   = 1u << 3,   // reported interval shares an endpoint with the cell it was created as
    kResPointAtEntry  = 1u << 4,   // returned by the isPoint() early exit
};

template <int D>
struct ResultRec {
    double root[D];                // getFinalPoint()
    double box[D][2];              // getFinalInterval()
    double comb_iv[D][2];          // getIntervalForCombining(): tracked interval used for overlap tests
    double iv_lo[D];               // tracked interval lower corner at emission (final-step de-dup key)
    double cell_iv[D][2];          // tracked interval of the box when it entered the kernel
    double next_mid[D];
    int32_t system, parent_node, collector, level;
    uint32_t flags; uint32_t pad;
    uint64_t path_hi, path_lo;
};

// One subdividing (non final-step) box; kept so the host can find lowest common
// ancestors when exterior boxes have to be merged (ChebyshevSubdivisionSolver.py:1318-1385).
template <int D>
struct NodeRec {
    double  iv[D][2];              // tracked interval at entry
    double  next_mid[D];
    int32_t system, parent_node, level, nsplit;
};

// Device-side counters of one level launch (zeroed by the host before the launch).
struct LevelCounters {
    unsigned long long next_box;        // work-distribution cursor
    unsigned long long n_children;      // records written to the next level
    unsigned long long data_used;       // doubles written to the next level's pool
    unsigned long long n_results;
    unsigned long long n_nodes;
    unsigned long long boxes;           // invocations of the reference's solvePolyRecursive
    unsigned long long zooms;
    unsigned long long subdivisions;
    unsigned long long discard_const, discard_quad, discard_zoom;
    unsigned long long entry_coeffs;    // sum of coefficients of boxes at entry (algorithmic bytes / 16)
    unsigned long long restrict_macs;   // multiply-adds performed by restrictions (triangular count)
    unsigned long long max_child_doubles;   // largest child box in the next level
    unsigned long long max_child_extent;    // largest axis extent in the next level
    unsigned long long max_level;
    unsigned long long overflow;        // non-zero: an output buffer was too small (host must retry)
    unsigned long long max_child_tensor;    // largest single tensor in the next level
    unsigned long long no_axis;             // subdivisions with no axis to halve along (the reference raises: :1010)
    unsigned long long reserved;
};

template <int D>
struct LevelArgs {
    const BoxRec<D>* recs_in; const double* data_in; unsigned long long n_in;
    BoxRec<D>* recs_out; double* data_out; unsigned long long cap_recs_out, cap_data_out;
    ResultRec<D>* results; unsigned long long cap_results;   // this launch appends at results[0..cap)
    NodeRec<D>* nodes; unsigned long long cap_nodes;
    unsigned long long result_base, node_base;               // global index of results[0] / nodes[0]
    LevelCounters* ctr;
    double* gws; unsigned long long gws_stride;   // per-CTA global workspace (doubles) when shared memory is too small
    unsigned long long ws_doubles;                // workspace capacity per CTA (doubles)
    SolverParams prm;
};

// ---- phase-split pipeline (yr_phases.h) ------------------------------------------------------
// Per-box working state of one level.  It lives in HBM only while the level is being processed:
// the scalar kernel (thread per box) and the tensor kernels (warp / CTA per box) hand a box back
// and forth through it.
constexpr int low_count(int D) { return (D + 1) * (D + 2) / 2; }   // coefficients of total degree <= 2

enum BoxPhase : int32_t {
    kPhNew = 0,        // sums + low-order coefficients are ready: node entry checks come next
    kPhTrimmed,        // trimMs done: initialise the zoom loop, first BILS
    kPhZoomed,         // a restriction was applied: zoom-loop bookkeeping, next BILS or finish
    kPhDone            // discarded, emitted, or queued for subdivision
};
enum TensorOp : int32_t { kOpNone = 0, kOpPrep, kOpTrim, kOpRestrict };

template <int D>
struct BoxWork {
    double sumabs[D];                 // sum|M_p| of the current tensors
    double low[D][low_count(D)];      // [c0 | linear[D] | pure T2[D] | cross pairs (a<b)] per polynomial
    double entry_iv[D][2];            // originalInterval.getIntervalForCombining() of the current invocation
    double cell_iv[D][2];             // tracked interval when the box entered the level
    double last_size[D];
    double alpha[D], beta[D];         // pending restriction (kOpRestrict)
    int32_t phase, op;
    int32_t zoom_count, changed, should_stop, final_split;
    int32_t node_level, result_index;
};

// Cursors and list lengths of one level's round loop (device memory, 16 x uint64).
struct RoundCtl {
    unsigned long long cursor;        // work-distribution cursor of the running kernel
    unsigned long long n_next;        // boxes queued for the next tensor step
    unsigned long long n_sub;         // boxes queued for subdivision
    unsigned long long max_total, max_tensor, max_extent;          // largest box queued for the next tensor step
    unsigned long long sub_max_total, sub_max_tensor, sub_max_extent;   // largest box queued for subdivision
    unsigned long long reserved[7];
};

template <int D>
struct PhaseArgs {
    LevelArgs<D> lv;                  // NB this pipeline updates lv.recs_in in place while a level is processed
    BoxWork<D>* work;                 // [lv.n_in]
    const unsigned int* list_in;      // boxes of the current step (nullptr: 0 .. n_list-1)
    unsigned long long n_list;
    unsigned int* list_next;          // boxes that need another tensor step
    unsigned int* list_sub;           // boxes to subdivide once the round loop is over
    RoundCtl* ctl;
    int first_round;                  // 1: every box of the level gets kOpPrep and a fresh BoxWork
};

// Workspace need (doubles) of a box whose largest tensor has `max_size` coefficients,
// `total` coefficients overall and largest extent `nmax` -- the host sizes launches with it.
YR_HD unsigned long long workspace_doubles(int D, unsigned long long total, unsigned long long max_size,
                                           unsigned long long nmax, int exact, int subdivide = 1) {
    // subdivide = 0: a tensor step (prep / trim / restrict) needs neither the subdivision temporaries nor
    // the upper-half matrices
    unsigned long long tri = nmax * (nmax + 1) / 2;
    const unsigned long long nmat = subdivide ? 2ull * D : (unsigned long long)D;
    unsigned long long w = total + (subdivide ? (unsigned long long)(D > 1 ? D - 1 : 0) * max_size : 0ull) + nmat * tri;
    w += (2ull * D * nmax + 1) / 2 + 1;                       // rowsAfter (ints)
    if (exact) w += nmat * 6 * (nmax + 2);
    return w + 8;
}

}  // namespace yr
