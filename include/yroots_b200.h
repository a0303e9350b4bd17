/*
 * yroots_b200 -- C ABI of the B200-native Chebyshev subdivision solver.
 *
 * This is the drop-in boundary for the hot path of tylerjarvis/RootFinding (YRoots):
 *   yroots.solve -> ChebyshevApproximator.chebApproximate
 *                -> ChebyshevSubdivisionSolver.solveChebyshevSubdivision
 * The reference is pure Python + numba and has no FFI of its own; the functions below are
 * what a binding for that path would call (see INTEGRATION.md for the ctypes stub a
 * maintainer would add to yroots/ChebyshevSubdivisionSolver.py).  All pointers are DEVICE
 * pointers unless stated otherwise; `stream` is a cudaStream_t passed as void* (NULL = the
 * legacy default stream).  All tensors are FP64, row-major.  No torch types cross this
 * boundary.  Every function returns 0 on success and a negative value on failure, with
 * yr_last_error() giving the reason.
 *
 * Reference interfaces replaced (file:line in /root/reference/yroots/):
 *   yr_init_boxes          ChebyshevSubdivisionSolver.py:1423-1434 (root box of a solve) and
 *                          :864-894 transformChebToInterval / :339-374 TransformChebInPlaceND
 *                          (batched restriction of whole systems to sub-boxes)
 *   yr_process_level       ChebyshevSubdivisionSolver.py:1177-1317 solvePolyRecursive, one
 *                          frontier level: constant check :1213, quadratic_check
 *                          (QuadraticCheck.py:25-687), trimMs :1131, zoomInOnIntervalIter :896 with
 *                          BoundingIntervalLinearSystem :628 and the restrictions :45-337,
 *                          getSubdivisionIntervals :1051
 *   yr_bils_batched        ChebyshevSubdivisionSolver.py:628-753 BoundingIntervalLinearSystem
 *   yr_quadcheck_batched   QuadraticCheck.py:25 quadratic_check
 *   yr_dct1_axis / yr_cheb_eval_axis / yr_abs_axis_mean
 *                          ChebyshevApproximator.py:28-89 interval_approximate_nd (grid evaluation of
 *                          Polynomial objects polynomial.py:317-342,503-528; dctn type 1 :78-82;
 *                          axis means of |coeff| :288)
 */
#ifndef YROOTS_B200_H
#define YROOTS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YR_ABI_VERSION 1
#define YR_MAX_DIM 6

/* Solver options; defaults are the reference's hard-coded constants (yr_default_params). */
typedef struct yr_params {
    int32_t exact;               /* SolverOptions.exact                      (ChebyshevSubdivisionSolver.py:35) */
    int32_t constant_check;      /* SolverOptions.constant_check             (:36) */
    int32_t low_dim_quad_check;  /* SolverOptions.low_dim_quadratic_check    (:37) */
    int32_t all_dim_quad_check;  /* SolverOptions.all_dim_quadratic_check    (:38) */
    int32_t max_zoom_count;      /* SolverOptions.maxZoomCount = 25          (:39) */
    int32_t use_fma;             /* 0 = separately rounded mul/add: restrictions bit-identical to the reference */
    int32_t trim;                /* 1 = run trimMs (:1232) */
    int32_t trim_policy;         /* which trim: 0 = trimMs (:1131-1171); 1, 2, 3 = trimMsOptimized1/2/3 of the reference's experiment
                                  * file tests/TrimMs optimization/ChebyshevSubdivisionSolverWithOptimizedTrimMs.py:1167-1271.
                                  * Policies 1-3 are served by yr_process_level only (yr_frontier_begin refuses them). */
    double  trim_rel_tol;        /* trimMs relApproxTol = 1e-3 (:1131) */
    double  trim_abs_tol;        /* trimMs absApproxTol = 0 */
} yr_params;

/* How a kernel is launched: the host sizes shared memory from the largest box of the level. */
typedef struct yr_launch {
    int32_t threads;             /* threads per CTA (multiple of 32) */
    int32_t ctas;                /* persistent CTAs */
    int32_t ws_in_smem;          /* 1: workspace in dynamic shared memory, 0: in gws */
    int32_t warp_per_box;        /* 0: CTA per box; 1: every warp of a CTA owns a box of its own (small boxes) */
    uint64_t ws_doubles;         /* workspace capacity per CTA, in doubles */
    double*  gws;                /* [ctas][gws_stride] global scratch when ws_in_smem == 0 */
    uint64_t gws_stride;
} yr_launch;

/* One frontier level (or chunk of a level). Record layouts: yr_record_layout(). */
typedef struct yr_level_desc {
    int32_t dim; int32_t reserved;
    const void*   recs_in;  const double* data_in;  uint64_t n_in;
    void*         recs_out; double*       data_out; uint64_t cap_recs_out; uint64_t cap_data_out;
    void*         results;  uint64_t cap_results;   /* this call appends at results[0 .. cap_results) */
    void*         nodes;    uint64_t cap_nodes;
    uint64_t      result_base; uint64_t node_base;  /* global indices of results[0] / nodes[0] (stored in links) */
    void*         counters;      /* yr_counters, zeroed by the caller before the call */
    yr_params     prm;
    yr_launch     launch;
    /* pipeline 1 (phase-split kernels): scratch of yr_level_work_bytes(dim, n_in) bytes that must stay
     * alive, untouched, until the level has succeeded; recs_in / data_in are updated in place.
     * resume_subdivide = 1 re-runs only the subdivision step after an overflow (larger *_out buffers). */
    void*         work;          uint64_t work_bytes;
    int32_t       pipeline;      /* must be 1: phase-split kernels (tensor step / scalar step rounds, then the subdivision step) */
    int32_t       resume_subdivide;
    /* pipeline 1 plans its own launches from the size of the largest box of the level: */
    uint64_t      max_box_doubles; uint64_t max_tensor_doubles; uint64_t max_extent;
} yr_level_desc;

/* Level-0 boxes: one per job = (system, optional restriction). */
typedef struct yr_init_desc {
    int32_t dim; int32_t reserved;
    const double*  coeffs;       /* all tensors of all systems */
    const int64_t* coeff_off;    /* [nsys][dim] offset (in doubles) of tensor p of system s */
    const int32_t* shapes;       /* [nsys][dim][dim] extents */
    const double*  errs;         /* [nsys][dim] approximation errors */
    const int32_t* job_system;   /* [njobs] */
    const int32_t* job_restrict; /* [njobs] 0 = take the system as is, 1 = restrict by alpha/beta */
    const double*  job_alpha;    /* [njobs][dim] */
    const double*  job_beta;     /* [njobs][dim] */
    const double*  job_iv;       /* [njobs][dim][2] interval of the new box in the solve's unit cube */
    const int32_t* job_level;    /* [njobs] */
    const double*  job_next_mid; /* [njobs][dim] */
    const int32_t* job_parent;   /* [njobs] */
    uint64_t       njobs;
    void*          recs_out; double* data_out; uint64_t cap_recs_out; uint64_t cap_data_out;
    void*          counters;
    yr_params      prm;
    yr_launch      launch;
} yr_init_desc;

/* Counters written by yr_process_level / yr_init_boxes (all uint64). */
typedef struct yr_counters {
    uint64_t next_box, n_children, data_used, n_results, n_nodes;
    uint64_t boxes, zooms, subdivisions, discard_const, discard_quad, discard_zoom;
    uint64_t entry_coeffs, restrict_macs, max_child_doubles, max_child_extent, max_level, overflow;
    uint64_t max_child_tensor;
    uint64_t no_axis;            /* boxes whose subdivision had no axis left (point-sized box at level <= 5): the reference's
                                  * getInverseOrder raises IndexError there (ChebyshevSubdivisionSolver.py:1010); the solve is void */
    uint64_t reserved;
} yr_counters;

int         yr_abi_version(void);
const char* yr_last_error(void);
/* class of the last failure on this thread: 0 none, 1 generic, 2 the DEVICE ran out of memory (a caller may retry with a
 * smaller batch), 3 a box had no axis left to subdivide along -- the input on which the reference raises IndexError
 * (getInverseOrder, ChebyshevSubdivisionSolver.py:1010) */
int         yr_last_error_code(void);
const char* yr_build_info(void);                 /* "cuda sm_100a ..." */
void        yr_default_params(yr_params* out);

/* Layout of the device records, so hosts do not hard-code it.  kind: 0 box, 1 result, 2 node.
 * Writes up to cap triples (field id, byte offset, byte size) and returns the record size. */
int64_t     yr_record_layout(int32_t dim, int32_t kind, int64_t* triples, int32_t cap);

/* Workspace (doubles) a box needs: total coefficients, largest tensor, largest extent. */
uint64_t    yr_workspace_doubles(int32_t dim, uint64_t total, uint64_t max_tensor, uint64_t max_extent, int32_t exact);
/* Scratch bytes yr_process_level needs for a level of n boxes when desc.pipeline == 1. */
uint64_t    yr_level_work_bytes(int32_t dim, uint64_t n_boxes);
/* Static shared memory (bytes) the solver kernels use besides the workspace, per CTA. */
int64_t     yr_static_smem_bytes(int32_t dim);
/* Largest dynamic shared memory a CTA may use on the current device (bytes); host pointers. */
int         yr_device_limits(int32_t* sm_count, int64_t* max_smem_per_cta, int64_t* smem_per_sm);

int yr_init_boxes(const yr_init_desc* d, void* stream);
int yr_process_level(const yr_level_desc* d, void* stream);

/* --- 2-D frontier pipeline ------------------------------------------------------------------
 * yr_frontier_solve runs the whole of solveChebyshevSubdivision's search
 * (ChebyshevSubdivisionSolver.py:1177-1317 for every box of every system) for systems of two
 * polynomials in two variables, starting from the level-0 boxes yr_init_boxes wrote: rounds of
 * (tensor ops: restriction :864-894 / subdivision :1051-1129, a team per box) and (scalar step:
 * checks :1213-1224, trimMs :1131 adoption, BoundingIntervalLinearSystem :628, zoom bookkeeping
 * :1243-1252, thread per box) over one device-resident frontier that is not level-synchronous.
 * The library owns every buffer of the solve; result and node records (yr_record_layout kinds 1, 2)
 * stay on the device until yr_frontier_release and are read with yr_fetch.  exact = 0 only. */
typedef struct yr_frontier_desc {
    int32_t dim;                 /* 2, 3 or 4 (the resumable pieces below and yr_f2_tensor_ops: 2 only) */
    int32_t time_kernels;        /* 1: CUDA-event timing of the tensor / scalar launches (out->tensor_ms, scalar_ms) */
    const void*   recs_in;  const double* data_in;  uint64_t n_in;   /* level-0 boxes (yr_init_boxes) */
    uint64_t      max_extent;    /* largest extent of any of their tensors */
    uint64_t      result_base; uint64_t node_base;  /* global indices of the first record this solve appends */
    yr_params     prm;
} yr_frontier_desc;

/* one entry of yr_frontier_out.order for dim = 2; for dim = 3, 4 the entry has root[dim] (16 + 8 * dim bytes) */
typedef struct yr_sorted_rec { uint32_t index; int32_t system; uint32_t flags; int32_t collector; double root[2]; } yr_sorted_rec;

typedef struct yr_frontier_out {
    void*    results; uint64_t n_results;   /* device arrays owned by the library */
    void*    nodes;   uint64_t n_nodes;
    void*    order;                         /* yr_sorted_rec[n_results]: the records in the reference's reporting order
                                               (system, then depth-first path) with the fields a host needs for the
                                               common case, so that it can skip gathering from the full records */
    yr_counters counters;                   /* totals of the solve (boxes, zooms, subdivisions, discards, ...) */
    uint64_t rounds, launches, slots;
    double   tensor_ms, scalar_ms;          /* summed device time of the two kernel families (time_kernels = 1) */
} yr_frontier_out;

int yr_frontier_solve(const yr_frontier_desc* d, yr_frontier_out* out, void* stream);

/* The same solve in resumable pieces: yr_frontier_solve == begin + advance(0) + finish.  Between two advance calls the
 * host may move queued boxes to another process' frontier -- the box-level multi-GPU mode (one system's frontier
 * sharded over the GPUs of a node, SURVEY 8e: sibling boxes are independent, ChebyshevSubdivisionSolver.py:1314-1317).
 * One solve per process at a time.  n_in may be 0 (a rank that only receives boxes).
 *   advance   runs rounds until the queues are empty or max_rounds rounds have run (0: no limit)
 *   export    removes up to n_take queued boxes (tails of the pending queues, every (op, class) in proportion), packs
 *             BoxRec + Work + both tensors of each into xbuf (device memory, >= yr_frontier_export_bound(n_take) bytes);
 *             needs at least one completed round
 *   import    adopts the boxes of such a buffer: records into fresh slots, tensors into the input arena, queue entries
 *             for their pending tensor op.  parent_node / collector links keep the exporter's global indices
 *             (yr_frontier_desc.result_base / node_base must give every process its own index range). */
typedef struct yr_frontier_progress {
    uint64_t pending, pending_restrict, pending_subdivide;   /* boxes queued for the next round */
    uint64_t rounds;                                         /* rounds run so far */
    uint64_t boxes;                                          /* solvePolyRecursive invocations so far */
    uint64_t arena_doubles;                                  /* coefficient data (doubles) the live boxes occupy, upper bound */
} yr_frontier_progress;
int yr_frontier_begin(const yr_frontier_desc* d, void* stream);
int yr_frontier_advance(uint64_t max_rounds, yr_frontier_progress* p, void* stream);
uint64_t yr_frontier_export_bound(uint64_t n_take);
int yr_frontier_export(uint64_t n_take, void* xbuf, uint64_t cap_bytes, uint64_t* n_taken, uint64_t* bytes, void* stream);
int yr_frontier_import(const void* xbuf, uint64_t bytes, void* stream);
int yr_frontier_finish(yr_frontier_out* out, void* stream);
/* 1 when 2-D boxes with extents up to max_extent fit the pipeline's shared-memory budget (else use yr_process_level) */
int yr_frontier_fits(uint64_t max_extent);
/* the same for dim = 2, 3, 4 (0 for other dimensions): max_tensor = coefficients of the largest single tensor */
int yr_frontier_fits_dim(int32_t dim, uint64_t max_extent, uint64_t max_tensor);
/* Record tables of several processes' frontier solves (box-level multi-GPU), concatenated by the caller in rank order on the
 * device, become ONE library-owned output as if a single solve had produced them: links (rank << rank_shift | local index)
 * are remapped to positions in the concatenation (res_base[r] / node_base[r] = first position of rank r's records) and
 * out->order is computed by the same device sort as yr_frontier_finish.  The inputs stay the caller's.  2-D records. */
int yr_frontier_adopt(int32_t dim, const void* results, uint64_t n_results, const void* nodes, uint64_t n_nodes,
                      const uint64_t* res_base, const uint64_t* node_base, int32_t nranks, int32_t rank_shift,
                      yr_frontier_out* out, void* stream);
/* device -> device copy on `stream` (asynchronous): lets a host move library-owned record arrays into buffers of its own */
int yr_device_copy(void* dst, const void* src, uint64_t bytes, void* stream);
int yr_frontier_release(yr_frontier_out* out, void* stream);
/* the pipeline keeps its box tables, arenas and queues between solves (grow-only); this frees them */
int yr_frontier_trim(void* stream);
/* Compacts the records of SELECTED systems: every record of `table` (n records of rec_bytes bytes, the int32 system id at
 * byte system_off: result or node records of a frontier solve) whose system has sys_flags[system] != 0 is appended to
 * out_recs, its index in the table to out_index; *out_n (device counter, zeroed by the caller) counts them (records beyond
 * `cap` are counted, not written).  All pointers are device pointers.  The host side of a large batch uses it to bring
 * over full records only for the few systems that need the tree bookkeeping (collectors, merges). */
int yr_gather_by_system(const void* table, uint64_t n, uint32_t rec_bytes, uint32_t system_off, const uint8_t* sys_flags, uint64_t nsys,
                        uint32_t* out_index, void* out_recs, uint64_t cap, uint64_t* out_n, void* stream);
/* device -> host copy of `bytes` bytes on `stream`, synchronous (host_dst is a HOST pointer) */
int yr_fetch(void* host_dst, const void* dev_src, uint64_t bytes, void* stream);

/* The frontier pipeline's TENSOR kernels on caller-supplied 2-D boxes -- the batched forms of transformChebToInterval
 * (ChebyshevSubdivisionSolver.py:864-894, op = 1) and getSubdivisionIntervals (:1051-1129 with getSubdivisionDims :1013-1049,
 * op = 2) -- through exactly the launches yr_frontier_solve makes (same size classes, teams and code).  Every pointer is
 * a HOST pointer (an operator-level / test entry, not the solve path).  Box b has two tensors, row-major at coeffs +
 * off[b][p] with extents shape[b][p][2], and error bounds err[b][p].
 *   op 1: alpha / beta [n][2] per axis.  One output box per input box.
 *   op 2: iv [n][2][2] tracked interval, next_mid [n][2] (nextTransformPoints), level [n] (solverOptions.level of the
 *         calling solvePolyRecursive).  2^k output boxes per input box, children of one box consecutive, in the
 *         reference's order; out_parent names the input box.
 * Outputs per output box q and polynomial p: tensor (row-major packed at out_coeffs + out_off[q][p]), out_shape, out_err
 * (error after the transformation charges :831), out_sumabs (sum|M|), the outcome trimMs (:1131-1171) would have on it
 * (out_trim_shape, out_trim_err), and the box's tracked interval / nextTransformPoints (out_iv, out_next_mid). */
typedef struct yr_f2_ops_desc {
    int32_t op, flat;            /* op 1, flat = 0: inputs packed like level-0 boxes (even row strides possible); otherwise like arena tensors (odd row stride) */
    uint64_t n;
    const double* coeffs; const uint64_t* off; const int32_t* shape; const double* err;
    const double* alpha; const double* beta;
    const double* iv; const double* next_mid; const int32_t* level;
    yr_params prm;
    uint64_t cap_out, cap_coeffs; uint64_t* n_out;
    double* out_coeffs; uint64_t* out_off; int32_t* out_shape; double* out_err; double* out_sumabs;
    int32_t* out_trim_shape; double* out_trim_err; double* out_iv; double* out_next_mid; int32_t* out_parent;
} yr_f2_ops_desc;
int yr_f2_tensor_ops(const yr_f2_ops_desc* d, void* stream);

/* BoundingIntervalLinearSystem on n independent dim x dim problems (thread per problem).
 *   lin [n][dim][dim], c0 [n][dim], sumabs [n][dim], errs [n][dim], final_step [n]
 *   -> interval [n][dim][2], flags [n][3] = (changed, should_stop, throwOut) */
int yr_bils_batched(int32_t dim, int64_t n, const double* lin, const double* c0, const double* sumabs,
                    const double* errs, const int32_t* final_step, double* interval, int32_t* flags, void* stream);

/* quadratic_check on n tensors of identical dimension (thread per tensor).
 *   coeffs: packed tensors, off [n], shapes [n][dim], tol [n], nd_check: force the n-D path
 *   -> out [n] (1 = no root possible) */
int yr_quadcheck_batched(int32_t dim, int64_t n, const double* coeffs, const int64_t* off, const int32_t* shapes,
                         const double* tol, int32_t nd_check, int32_t* out, void* stream);

/* --- approximator (ChebyshevApproximator.py) ------------------------------------------- */
/* values [outer][n+1][inner] -> coeffs along the middle axis: DCT-I with the reference's
 * 1/deg scaling and halved end planes (interval_approximate_nd :78-82), batched. */
int yr_dct1_axis(const double* values, double* coeffs, int64_t outer, int64_t npts, int64_t inner, void* stream);
/* Evaluate a Chebyshev (basis 0) or power (basis 1) coefficient tensor along one axis at npts
 * points x: in [outer][ncoef][inner] -> out [outer][npts][inner]. */
int yr_poly_eval_axis(const double* coef, const double* x, double* out, int64_t outer, int64_t ncoef,
                      int64_t npts, int64_t inner, int32_t basis, void* stream);
/* mean of |t| over all axes but one: in [outer][n][inner] -> out [n]; also max|t| -> *supnorm. */
int yr_abs_axis_mean(const double* t, double* out, double* supnorm, int64_t outer, int64_t n, int64_t inner, void* stream);

/* out[o][i][in] = sum_{k >= i} P[i][k] * coef[o][k][in] (P row-major n x n, upper triangular part used), accumulated from
 * zero in ascending k with separately rounded multiply and add: MultiPower.to_cheb along one axis (polynomial.py:587-640,
 * P[i][k] = T_i coefficient of x^k), bit-identical to the reference's loop. */
int yr_upper_axis_apply(const double* coef, const double* P, double* out, int64_t outer, int64_t n, int64_t inner, void* stream);

/* One probe of getChebyshevDegrees' doubling search decided on the device (ChebyshevApproximator.py:228-311): the host
 * reads back this 64-byte struct instead of the coefficient tensors.
 *   mode 0  first probe of an iteration (degree guess g along `axis`): supnorm = max|values| (:72), tol = abs_tol +
 *           supnorm * rel_tol (:290), started = startedConverging(axis means of |coeff|, tol) (:91-106, :288)
 *   mode 1  second probe (degree 2g+1): supnorm of ITS values, tol from max(prev_supnorm, supnorm) (:300), converged =
 *           hasConverged(prev_coeff, coeff, tol) (:108-129), and (degree, eps, rho) = getFinalDegree(axis means, tol) (:131-172)
 * coeff / prev_coeff: full coefficient tensors on the device, row-major with extents shape[] / prev_shape[]; keep[] <= shape[]
 * is the sub-box the reference's sliced tensor covers (axes the function is constant along keep one plane, :57,:85).
 * scratch: keep[axis] + 128 doubles of device memory. */
typedef struct yr_decay_out { double supnorm, tol, started, converged, degree, eps, rho, reserved; } yr_decay_out;
int yr_decay_probe(int32_t mode, int32_t dim, int32_t axis, const double* coeff, const int64_t* shape, const int64_t* keep,
                   const double* prev_coeff, const int64_t* prev_shape, const int64_t* prev_keep, const double* values, int64_t nvalues,
                   double prev_supnorm, double rel_tol, double abs_tol, double* scratch, yr_decay_out* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YROOTS_B200_H */
