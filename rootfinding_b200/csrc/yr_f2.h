// 2-D frontier pipeline ("f2"): the Chebyshev subdivision solver for systems of two polynomials in two
// variables, organised for throughput on a B200.
//
// The whole search (every solvePolyRecursive invocation of every system of a batch) runs as a sequence of
// ROUNDS over one device-resident frontier.  A round is
//
//   tensor ops   team (warp, or CTA of 4 / 8 warps) per box, one launch per (op, size class):
//                  RESTRICT   transformChebToInterval (:864-894): build C(alpha,beta) per axis, restrict both
//                             tensors along both axes
//                  SUBDIVIDE  getSubdivisionIntervals (:1051-1129): halve both tensors along the split axes with
//                             the level-independent matrices C(1/2 +- m/2, ...) kept in a global table
//                every produced tensor leaves the team with its sum|M|, its coefficients of total degree <= 2
//                and the outcome trimMs (:1131-1171) would have on it, so that ...
//   scalar step  thread per box: ... isPoint, constant check (:1213), quadratic check (QuadraticCheck.py:34-182),
//                trimMs adoption, BoundingIntervalLinearSystem (:628-753), addTransform (:420-454), the zoom-loop
//                bookkeeping (:1243-1252), final-step re-entry (:1264), result / node records, and the choice of
//                the box's next tensor op never touch a tensor.
//
// Boxes are not level-synchronous: a box that needs another zoom simply appears in the next round's RESTRICT
// list next to boxes of other levels and other systems, so the number of rounds is the length of the longest
// dependency chain, not (levels x longest zoom loop).  Tensors ping-pong between two arenas (every live box is
// rewritten by exactly one tensor op per round); box records live in stable slots.
//
// Layout of a tensor in an arena and in shared memory: row-major with an ODD row stride ld >= n1, so that both
// restriction passes read shared memory without bank conflicts and global <-> shared copies are flat.
// Restriction matrices are stored in 4-row panels: panel b holds C[4b..4b+3][k], k = 4b..n-1, as [k][4] -- one
// item (fiber f, row block b) loads a source element once and feeds four accumulators from two 16-byte loads.
//
// Arithmetic: compiled with -fmad=false; every matrix entry is generated with the reference's expression order
// (:87-112) and every output accumulates its sources in ascending order, so restricted tensors are bit-identical
// to TransformChebInPlace1D (use_fma = 1 fuses the contraction's multiply-add).
//
// Written against a small `Team` interface (tid, nthreads, sync, sum, bcast) and backend hooks so that
// tests/emul can run the same source on pthreads.  The emulator is test infrastructure only.
#pragma once
#include "yr_types.h"

#ifndef YR_DEV
#if defined(__CUDACC__)
#define YR_DEV __device__
#define YR_DEVN __device__ __noinline__
#else
#define YR_DEV
#define YR_DEVN
#endif
#endif

// inlining policy of the big per-box helpers (code size vs call overhead; see profiles/)
#ifndef YR_F2_PASS
#define YR_F2_PASS YR_DEV
#endif
#ifndef YR_F2_BUILD
#define YR_F2_BUILD YR_DEV
#endif
#ifndef YR_F2_TRIM_QUICK
#define YR_F2_TRIM_QUICK 0
#endif
#ifndef YR_F2_PAIR_BUILD
#define YR_F2_PAIR_BUILD 0
#endif
#ifndef YR_F2_EPI
#define YR_F2_EPI YR_DEV
#endif
#ifndef YR_F2_SKIP_DEAD
#define YR_F2_SKIP_DEAD 1
#endif

namespace yr {
namespace f2 {

#if defined(__CUDACC__)
typedef double2 pair_t;
#else
struct alignas(16) pair_t { double x, y; };
#endif

// Size classes.  A class fixes the team (16 lanes, 32 lanes, CTA of 128 / 256) and bounds the shared memory a
// box may need; a launch sizes every team's slice by the largest need actually queued in that class, so finer
// classes mean more resident teams per SM for the typical box.
constexpr int kOps = 2;
enum Op : int32_t { kTNone = 0, kTRestrict = 1, kTSubdivide = 2 };
YR_HD int op_slot(int op) { return op == kTSubdivide ? 1 : 0; }

constexpr int kClasses = 15;
constexpr int kWarpClasses = 9;             // classes below this run the warp-team code
constexpr int kLane16Classes = 5;           // classes below this use two 16-lane teams per warp
constexpr int kCta128Classes = 13;          // CTA classes below this run 128 threads per box, the rest 256
YR_HD unsigned long long class_bound(int c) {               // no table: an indexed local array would live on the stack
    switch (c) {
        case 0: return 384ull; case 1: return 512ull; case 2: return 640ull; case 3: return 768ull; case 4: return 1024ull;
        case 5: return 1536ull; case 6: return 2048ull; case 7: return 3072ull; case 8: return 4096ull;
        case 9: return 2048ull; case 10: return 3072ull; case 11: return 4608ull; case 12: return 7168ull;
        case 13: return 12288ull;
        default: return ~0ull;
    }
}
YR_HD int warp_class(unsigned long long need) {             // smallest warp class whose bound holds `need` (<= 4096)
    return (need > 384ull) + (need > 512ull) + (need > 640ull) + (need > 768ull) + (need > 1024ull) + (need > 1536ull) +
           (need > 2048ull) + (need > 3072ull);
}
YR_HD int cta_class(unsigned long long need, unsigned long long split) {   // `split`: largest need a 128-thread CTA takes
    if (need <= split) return kWarpClasses + (need > 2048ull) + (need > 3072ull) + (need > 4608ull);
    return kCta128Classes + (need > 12288ull);
}
YR_HD int class_threads(int cls) { return cls < kLane16Classes ? 16 : (cls < kWarpClasses ? 32 : (cls < kCta128Classes ? 128 : 256)); }
constexpr unsigned long long kWarp32Max = 4096ull;    // largest need a warp team can take
constexpr unsigned long long kMaxNeed = 28800ull;     // doubles of dynamic shared memory a CTA may use (225 KiB of 227)

// Launch-shape policy (host-tunable; defaults from the sweeps in profiles/r02_experiments.txt and profiles/r02_fd.txt): the largest warp-team need per op
// (boxes above it run one CTA per box) and the largest 128-thread CTA need.
struct TeamPolicy { unsigned long long warp_max[2]; unsigned long long cta128_max; };

YR_HD int odd_ld(int n1) { return n1 | 1; }
YR_HD int even_up(int x) { return (x + 1) & ~1; }
YR_HD int tensor_doubles(int n0, int ld) { return even_up(n0 * ld); }

// ---- 4-row panels ------------------------------------------------------------------------------
YR_HD int panel_off(int b, int nt) { return 4 * b * nt - 8 * b * (b - 1); }
YR_HD int panel_doubles(int nt) { const int B = (nt + 3) >> 2; return panel_off(B, nt); }
YR_HD int panel_index(int i, int k, int nt) { const int b = i >> 2; return panel_off(b, nt) + ((k - 4 * b) << 2) + (i & 3); }

YR_HD unsigned long long need_restrict(int T, int n0max, int n1max) {          // CTA teams: [A][B][X][C0][C1][rows]
    return 3ull * T + panel_doubles(n0max) + panel_doubles(n1max) + (unsigned)((n0max + n1max + 1) / 2) + 2;
}
YR_HD unsigned long long need_subdivide(int T) { return 4ull * T; }              // CTA teams: [S][L][U][X]
YR_HD unsigned long long need_restrict_w(int T, int n0max, int n1max) {        // warp teams: [A][C0][C1][rows]
    return (unsigned long long)T + panel_doubles(n0max) + panel_doubles(n1max) + (unsigned)((n0max + n1max + 1) / 2) + 2;
}
YR_HD unsigned long long need_subdivide_w(int T) { return 3ull * T; }            // warp teams: [S][L][X]
YR_HD void pick_team(unsigned long long need_w, unsigned long long need_cta, int coarse, const TeamPolicy& pol, int slot,
                     unsigned long long& need, int& cls) {
    if (need_w <= pol.warp_max[slot]) {
        need = need_w; cls = warp_class(need_w);
        // a small round is launch-latency bound: one class per team type, fewer launches
        if (coarse) cls = cls < kLane16Classes ? kLane16Classes - 1 : kWarpClasses - 1;
    } else {
        need = need_cta; cls = cta_class(need_cta, pol.cta128_max);
        if (coarse) cls = cls < kCta128Classes ? kCta128Classes - 1 : kClasses - 1;
    }
}

// Per-box state next to BoxRec<2> (same slot index).
struct Work {
    unsigned long long off[2];   // first double of tensor p in the arena the box currently lives in
    int32_t ld[2];               // its row stride (odd once the box has been through a tensor op)
    int32_t tshape[2][2];        // extents trimMs would leave              } valid for the tensors as the last
    double  tsumabs[2], terr[2]; // sum|M_p| and error after that trim      } tensor op wrote them (trim_valid)
    double  sumabs[2];           // sum|M_p| of the current view
    double  low[2][6];           // M00, M10, M01, M20, M11, M02 (0 beyond the extents): quad_check_2d order
    double  alpha[2], beta[2];   // pending restriction
    double  entry_iv[2][2];      // originalInterval.getIntervalForCombining() of the current invocation
    double  cell_iv[2][2];       // tracked interval when the box was created
    double  last_size[2];
    int32_t op, fresh, copy_only, trim_valid;
    int32_t zoom_count, changed, should_stop, final_split;
    int32_t node_level, result_index, node_index, nsplit;
    int32_t order[2][2];         // per polynomial: axes to halve, in order
};

// Device control block; the host reads it once per round.
struct Ctl {
    unsigned long long n_list[kOps][kClasses];      // boxes queued for the next round
    unsigned long long need_max[kOps][kClasses];    // largest per-team need (doubles) among them
    unsigned long long ext_max[kClasses];           // largest extent among queued subdivisions
    unsigned long long out_bound;                   // doubles the next round's tensor ops may write
    unsigned long long data_used;                   // doubles written to the current output arena
    unsigned long long cursor[kOps][kClasses];      // this round's tensor launches: next queue position to hand out
    unsigned long long n_boxes, n_results, n_nodes; // slots in use
    unsigned long long overflow, bad_mid, no_axis;
    unsigned long long boxes, zooms, subdivisions, dconst, dquad, dzoom, entry_coeffs, restrict_macs, max_level;
    unsigned long long reserved[4];
};
constexpr int kCtlWords = sizeof(Ctl) / 8;

struct Tables {                // level-independent subdivision matrices, panels for extent `nt`
    const double* mat[2][2];   // [mid: 0 -> m = 0, 1 -> m = kFirstSplit][half: 0 lower, 1 upper]
    const int* rows_after[2][2];
    int nt;
};

struct Args {
    BoxRec<2>* recs; Work* work;
    const double* data_in; double* data_out; unsigned long long cap_data_out;
    ResultRec<2>* results; unsigned long long cap_results, result_base;
    NodeRec<2>* nodes; unsigned long long cap_nodes, node_base;
    unsigned int* lists[kOps][kClasses];            // queues being filled for the next round
    unsigned long long cap_boxes;
    int coarse;                                     // 1: the next round is small, queue into one class per team type
    TeamPolicy pol;
    Ctl* ctl;
    Tables tab;
    SolverParams prm;
};

// ---- restriction matrix C(alpha, beta) in panel form ---------------------------------------------
// Column by column (the three-term recurrence couples neighbouring rows: columns sequential, rows parallel),
// with rows_after[k] = the reference's `maxRow` after column k (:106-112).  Collective over the team.
template <class Team>
YR_DEVN void build_panels(Team& tm, double* Cp, int* rows_after, int nt, double alpha, double beta) {
    const int P = panel_doubles(nt);
    for (int e = tm.tid; e < P; e += tm.nthreads) Cp[e] = 0.0;
    tm.sync();
    if (tm.tid == 0) {
        Cp[panel_index(0, 0, nt)] = 1.0; rows_after[0] = 1;
        if (nt >= 2) { Cp[panel_index(0, 1, nt)] = beta; Cp[panel_index(1, 1, nt)] = alpha; rows_after[1] = 2; }
    }
    tm.sync();
    const double beta2 = 2 * beta;
    for (int k = 2; k < nt; ++k) {
        const int rows = rows_after[k - 1];
        for (int i = tm.tid; i <= rows; i += tm.nthreads) {
            // entries beyond a column's length are zero, which reproduces the reference's edge formulas exactly
            const double cm1 = (i >= 1) ? Cp[panel_index(i - 1, k - 1, nt)] : 0.0;
            const double ci = (i <= k - 1) ? Cp[panel_index(i, k - 1, nt)] : 0.0;
            const double cp1 = (i + 1 <= k - 1) ? Cp[panel_index(i + 1, k - 1, nt)] : 0.0;
            const double pi = (i <= k - 2) ? Cp[panel_index(i, k - 2, nt)] : 0.0;
            double t;
            if (i == 0) t = alpha * cp1;
            else if (i == 1) t = alpha * (2 * cm1 + cp1);
            else t = alpha * (cm1 + cp1);
            double val = (-pi + t) + beta2 * ci;
            if (i == rows) {                                  // the reference's finalVal
                val = alpha * cm1;
                int nr = rows;
                if (fabs(val) > 1e-16) nr = rows + 1; else val = 0.0;
                rows_after[k] = nr;
            }
            Cp[panel_index(i, k, nt)] = val;
        }
        tm.sync();
    }
}

// The same matrix built by ONE warp: lane i keeps row i of the two live columns in registers, neighbours come
// from shuffles, the lane walks its own panel row with a running pointer.  `seg` lanes form a segment that builds
// one matrix (seg = 32: one matrix per warp; seg = 16: lanes 0-15 build matrix A while lanes 16-31 build matrix B,
// both loops fused).  Same expressions, same order, same truncation rule as build_panels -> bit-identical.
// Requires nt <= seg for every matrix built.  `W` provides lane, wshfl_up/down/idx(v, delta|src, width), wsync().
struct PanelJob { double* Cp; int* rows_after; int nt; double alpha, beta; };

template <class W>
YR_F2_BUILD void build_panels_warp(W& w, const PanelJob& ja, const PanelJob& jb, int seg) {
    const int lane = w.lane;
    const bool second = (seg == 16) && (lane >= 16);
    const PanelJob& j = second ? jb : ja;
    const int i = (seg == 16) ? (lane & 15) : lane;
    const int nt = j.nt;
    const int kmax = (seg == 16) ? (ja.nt > jb.nt ? ja.nt : jb.nt) : ja.nt;
    const int B4 = ((nt + 3) >> 2) << 2;                       // rows that exist in the panels (zero rows pad the last one)
    const int b = i >> 2;
    double* ptr = j.Cp + panel_off(b, nt) + (i & 3);           // entry (i, k) lives at ptr[(k - 4b) * 4]
    const double alpha = j.alpha, beta = j.beta, beta2 = 2 * beta;
    int* const rows_after = j.rows_after;                      // (a register copy: the job lives in local memory and the stores below may alias it)
    const bool live = i < B4 && nt > 0;
    double prev2 = (i == 0) ? 1.0 : 0.0;                       // column 0
    double prev = (i == 0) ? beta : (i == 1 ? alpha : 0.0);    // column 1
    if (live && b == 0) { ptr[0] = prev2; if (nt >= 2) ptr[4] = prev; }
    if (i == 0 && nt > 0) { rows_after[0] = 1; if (nt >= 2) rows_after[1] = 2; }
    int rows = 2;
    const int segshift = (seg == 16 && w.lanes == 32) ? (lane & 16) : 0;
    for (int k = 2; k < kmax; ++k) {
        double cm1 = w.wshfl_up(prev, 1, seg), cp1 = w.wshfl_down(prev, 1, seg);
        if (i == 0) cm1 = 0.0;
        if (i + 1 > k - 1 || i + 1 >= seg) cp1 = 0.0;
        // every lane evaluates both candidates; selects instead of branches (same expressions as build_panels)
        const double t = alpha * ((i == 1 ? 2 * cm1 : cm1) + cp1);
        const double inner = (-prev2 + t) + beta2 * prev;
        const double fin = alpha * cm1;                          // the reference's finalVal, lane `rows` only
        const bool last = (i == rows);
        const bool g = last && (fabs(fin) > 1e-16);
        const double val = last ? (g ? fin : 0.0) : (i < rows ? inner : 0.0);
        const int grow = ((w.wballot(g) >> segshift) & (seg == 16 ? 0xffffu : 0xffffffffu)) != 0u;
        if (k < nt) {
            if (live && k >= 4 * b) ptr[(k - 4 * b) << 2] = val;
            prev2 = prev; prev = val;
            rows += grow;
            if (i == (k & (seg - 1))) rows_after[k] = rows;
        }
    }
    w.wsync();
}

// Two matrices at once, every lane holding row i of BOTH (16-lane teams: their two builds are two independent
// dependency chains of shuffles and separately rounded multiply-adds; interleaved they hide each other's latency).
// Same expressions, order and truncation rule per matrix as build_panels_warp -> bit-identical.  nt <= lanes, both > 0.
template <class W>
YR_F2_BUILD void build_panels_pair(W& w, const PanelJob& ja, const PanelJob& jb) {
    const int seg = w.lanes;
    const int i = w.lane;
    const int nta = ja.nt, ntb = jb.nt;
    const int kmax = nta > ntb ? nta : ntb;
    const int b = i >> 2;
    double* pa = ja.Cp + panel_off(b, nta) + (i & 3);
    double* pb = jb.Cp + panel_off(b, ntb) + (i & 3);
    const bool livea = i < (((nta + 3) >> 2) << 2), liveb = i < (((ntb + 3) >> 2) << 2);
    const double ala = ja.alpha, be2a = 2 * ja.beta, alb = jb.alpha, be2b = 2 * jb.beta;
    int* const ra = ja.rows_after; int* const rb = jb.rows_after;      // (register copies, see build_panels_warp)
    double p2a = (i == 0) ? 1.0 : 0.0, p2b = p2a;
    double p1a = (i == 0) ? ja.beta : (i == 1 ? ala : 0.0), p1b = (i == 0) ? jb.beta : (i == 1 ? alb : 0.0);
    if (b == 0) {
        if (livea) { pa[0] = p2a; if (nta >= 2) pa[4] = p1a; }
        if (liveb) { pb[0] = p2b; if (ntb >= 2) pb[4] = p1b; }
    }
    if (i == 0) {
        ra[0] = 1; if (nta >= 2) ra[1] = 2;
        rb[0] = 1; if (ntb >= 2) rb[1] = 2;
    }
    int rowsa = 2, rowsb = 2;
    for (int k = 2; k < kmax; ++k) {
        double cm1a = w.wshfl_up(p1a, 1, seg), cp1a = w.wshfl_down(p1a, 1, seg);
        double cm1b = w.wshfl_up(p1b, 1, seg), cp1b = w.wshfl_down(p1b, 1, seg);
        if (i == 0) { cm1a = 0.0; cm1b = 0.0; }
        if (i + 1 > k - 1 || i + 1 >= seg) { cp1a = 0.0; cp1b = 0.0; }
        const double ta = ala * ((i == 1 ? 2 * cm1a : cm1a) + cp1a), tb = alb * ((i == 1 ? 2 * cm1b : cm1b) + cp1b);
        const double ina = (-p2a + ta) + be2a * p1a, inb = (-p2b + tb) + be2b * p1b;
        const double fa = ala * cm1a, fb = alb * cm1b;
        const bool lasta = (i == rowsa), lastb = (i == rowsb);
        const bool ga = lasta && (fabs(fa) > 1e-16), gb = lastb && (fabs(fb) > 1e-16);
        const double va = lasta ? (ga ? fa : 0.0) : (i < rowsa ? ina : 0.0);
        const double vb = lastb ? (gb ? fb : 0.0) : (i < rowsb ? inb : 0.0);
        const int growa = w.wballot(ga) != 0u, growb = w.wballot(gb) != 0u;
        if (k < nta) {
            if (livea && k >= 4 * b) pa[(k - 4 * b) << 2] = va;
            p2a = p1a; p1a = va; rowsa += growa;
            if (i == (k & (seg - 1))) ra[k] = rowsa;
        }
        if (k < ntb) {
            if (liveb && k >= 4 * b) pb[(k - 4 * b) << 2] = vb;
            p2b = p1b; p1b = vb; rowsb += growb;
            if (i == (k & (seg - 1))) rb[k] = rowsb;
        }
    }
    w.wsync();
}

// Team-level dispatcher: builds up to two matrices (either may be skipped with nt = 0).
template <class Team>
YR_F2_BUILD void build_two(Team& tm, const PanelJob& j0, const PanelJob& j1) {
    const int big = j0.nt > j1.nt ? j0.nt : j1.nt;
    if (Team::kShuffle && (tm.lanes == 32 || tm.lanes == 16) && big <= tm.lanes) {
        if (tm.warps >= 2) {                                      // one warp per matrix
            if (tm.warp == 0 && j0.nt > 0) build_panels_warp(tm, j0, j0, 32);
            if (tm.warp == 1 && j1.nt > 0) build_panels_warp(tm, j1, j1, 32);
        } else if (tm.lanes == 32 && big <= 16 && j0.nt > 0 && j1.nt > 0) build_panels_warp(tm, j0, j1, 16);
        else if (YR_F2_PAIR_BUILD && tm.lanes == 16 && j0.nt > 0 && j1.nt > 0) build_panels_pair(tm, j0, j1);
        else {
            const PanelJob* jobs[2] = {&j0, &j1};
#pragma unroll 1
            for (int m = 0; m < 2; ++m) if (jobs[m]->nt > 0) build_panels_warp(tm, *jobs[m], *jobs[m], tm.lanes);   // one copy of the loop body
        }
        tm.sync();
        return;
    }
    if (j0.nt > 0) build_panels(tm, j0.Cp, j0.rows_after, j0.nt, j0.alpha, j0.beta);
    if (j1.nt > 0) build_panels(tm, j1.Cp, j1.rows_after, j1.nt, j1.alpha, j1.beta);
}

// ---- one restriction pass ---------------------------------------------------------------------------
// src element (k, f) at src[k*sk + f*sf], k < n;  dst element (i, f) at dst[i*dk + f*df], i < r;  f < F.
// An item is (row block b, fiber f): four output rows from one sweep over source rows 4b .. n-1.
// Returns the team-wide sum|dst|.  src and dst must not overlap.
template <bool FMA, class Team>
YR_F2_PASS double restrict_pass(Team& tm, const double* src, int sk, int sf, double* dst, int dk, int df,
                             int n, int r, int F, const double* Cp, int nt, unsigned long long& macs) {
    const int B = (r + 3) >> 2;
    double asum = 0.0;
    // items in row-block-major order get shorter and shorter; waves alternate direction so that every thread's
    // total trip count is about the same (the pass ends in a team barrier)
    const int items = B * F, T = tm.nthreads;
    for (int w0 = 0, wave = 0; w0 < items; w0 += T, ++wave) {
        const int id = w0 + ((wave & 1) ? (T - 1 - tm.tid) : tm.tid);
        if (id >= items) continue;
        const int b = id / F, f = id - b * F;
        const int i0 = b << 2;
        const pair_t* cp = reinterpret_cast<const pair_t*>(Cp + panel_off(b, nt));      // 16-byte aligned: two loads per k
        const double* s = src + i0 * sk + f * sf;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 2
        for (int k = i0; k < n; ++k) {
            const double x = *s; s += sk;
            const pair_t c01 = cp[0], c23 = cp[1]; cp += 2;
            const double c0 = c01.x, c1 = c01.y, c2 = c23.x, c3 = c23.y;
            if (FMA) { a0 = fma(c0, x, a0); a1 = fma(c1, x, a1); a2 = fma(c2, x, a2); a3 = fma(c3, x, a3); }
            else { a0 = a0 + c0 * x; a1 = a1 + c1 * x; a2 = a2 + c2 * x; a3 = a3 + c3 * x; }
        }
        double* d = dst + i0 * dk + f * df;
        const int rem = r - i0;
        d[0] = a0; asum += fabs(a0);
        if (rem > 1) { d[dk] = a1; asum += fabs(a1); }
        if (rem > 2) { d[2 * dk] = a2; asum += fabs(a2); }
        if (rem > 3) { d[3 * dk] = a3; asum += fabs(a3); }
    }
    if (tm.tid == 0) macs += (unsigned long long)(unsigned)(F * (r * n - r * (r - 1) / 2));     // triangular count
    return tm.sum_sync(asum);            // also the barrier that publishes dst
}

template <class Team>
YR_DEV void copy_flat(Team& tm, double* dst, const double* src, int n_even) {
    // both 16-byte aligned, n_even even
    const pair_t* s = reinterpret_cast<const pair_t*>(src);
    pair_t* d = reinterpret_cast<pair_t*>(dst);
    for (int e = tm.tid; e < (n_even >> 1); e += tm.nthreads) d[e] = s[e];
}

template <class Team>
YR_DEVN double view_abs_sum(Team& tm, const double* M, int n0, int n1, int ld) {
    double s = 0.0;
    if (n1 >= 12) {
        // lanes along rows
        const int per = tm.nthreads;
        for (int e = tm.tid; e < n0 * n1; e += per) { const int i = e / n1; s += fabs(M[i * ld + (e - i * n1)]); }
    } else {
        for (int i = tm.tid; i < n0; i += tm.nthreads) { const double* row = M + i * ld; for (int j = 0; j < n1; ++j) s += fabs(row[j]); }
    }
    return tm.sum(s);
}

// What every produced tensor leaves behind: low-order coefficients and the outcome of trimMs (:1131-1171)
// for error `err`.  Runs on the team's LEAD WARP only (callers guard with tm.warp == 0); lane 0 == thread 0
// holds the results.  Children of a subdivision typically lose several trailing planes per axis, so groups of
// four lanes sum up to lanes/4 candidate planes at once and the accept / stop scan then walks those sums.
struct Epilogue { double low[6]; int t0, t1; double tsum, terr; };

template <class Team>
YR_F2_EPI void tensor_epilogue(Team& tm, const double* M, int n0, int n1, int ld, double sumabs, double err,
                               const SolverParams& prm, Epilogue& ep) {
    ep.low[0] = M[0];
    ep.low[1] = n0 > 1 ? M[ld] : 0.0;
    ep.low[2] = n1 > 1 ? M[1] : 0.0;
    ep.low[3] = n0 > 2 ? M[2 * ld] : 0.0;
    ep.low[4] = (n0 > 1 && n1 > 1) ? M[ld + 1] : 0.0;
    ep.low[5] = n1 > 2 ? M[2] : 0.0;
    int t0 = n0, t1 = n1;
    double budget = prm.trim_abs_tol + err * prm.trim_rel_tol;
    double removed = 0.0;
    if (prm.trim) {
        const int gs = tm.lanes >= 4 ? 4 : tm.lanes;                 // lanes per plane
        const int G = tm.lanes / gs;                                  // planes per sweep
        const int g = tm.lane / gs, gl = tm.lane - g * gs;
        for (int axis = 0; axis < 2; ++axis) {
            for (;;) {
                const int t = axis ? t1 : t0;
                if (t <= 3) break;
                if (YR_F2_TRIM_QUICK) {
                    // One element of the last plane at or above the budget already rules the plane out (a rounded
                    // sum of non-negative terms is never below one of them): the common outcome "nothing to trim
                    // along this axis" costs one load and one vote instead of a sweep.  Same decision, exactly.
                    bool big = false;
                    if (axis == 0) { const double* row = M + (t0 - 1) * ld; for (int j = tm.lane; j < t1; j += tm.lanes) big = big || (fabs(row[j]) >= budget); }
                    else { const double* col = M + (t1 - 1); for (int i = tm.lane; i < t0; i += tm.lanes) big = big || (fabs(col[i * ld]) >= budget); }
                    if (tm.wany(big)) break;
                }
                const int np = (t - 3) < G ? (t - 3) : G;         // candidate planes t-1, t-2, ...
                double s = 0.0;
                if (g < np) {
                    if (axis == 0) { const double* row = M + (t0 - 1 - g) * ld; for (int j = gl; j < t1; j += gs) s += fabs(row[j]); }
                    else { const double* col = M + (t1 - 1 - g); for (int i = gl; i < t0; i += gs) s += fabs(col[i * ld]); }
                }
                s = tm.wsum_w(s, gs);
                int taken = 0;
                for (int q = 0; q < np; ++q) {
                    const double sq = tm.wshfl_idx_d(s, q * gs);
                    if (sq < budget) { budget -= sq; err += sq; removed += sq; ++taken; } else break;
                }
                if (axis) t1 -= taken; else t0 -= taken;
                if (taken < np) break;
            }
        }
    }
    ep.t0 = t0; ep.t1 = t1; ep.terr = err;
    ep.tsum = sumabs - removed;          // == sum over the trimmed view up to rounding (removed < 1e-3 * err)
}

YR_HD void store_epilogue(Work& w, int p, const Epilogue& ep, double sumabs) {
    w.sumabs[p] = sumabs;
    for (int q = 0; q < 6; ++q) w.low[p][q] = ep.low[q];
    w.tshape[p][0] = ep.t0; w.tshape[p][1] = ep.t1; w.tsumabs[p] = ep.tsum; w.terr[p] = ep.terr;
}

// ---- both halves of a subdivision in one sweep ---------------------------------------------------------
// For the split point m = 0 the two restriction matrices are mirror images, C_hi[i][k] = (-1)^(i+k) C_lo[i][k],
// EXACTLY (every step of the recurrence is sign-symmetric in IEEE arithmetic).  One product c*x therefore serves
// both halves -- lower += p, upper += +-p -- and because x +- p rounds like x + (+-p) the upper half is still
// bit-identical to a separate pass with C_hi: 4 DMUL + 8 DADD per source element instead of 8 + 8, half the
// matrix loads, one sweep over the source.  Sign for row block base i0 = 4b (even), row q, source k: (-1)^(q+k).
#define YR_F2_DUAL_STEP(SGN)                                                                                   \
    {                                                                                                          \
        const double x = *s; s += sk;                                                                          \
        const pair_t c01 = cp[0], c23 = cp[1]; cp += 2;                                                        \
        if (FMA) {                                                                                             \
            l0 = fma(c01.x, x, l0); l1 = fma(c01.y, x, l1); l2 = fma(c23.x, x, l2); l3 = fma(c23.y, x, l3);    \
            u0 = fma((SGN) * c01.x, x, u0); u1 = fma(-(SGN) * c01.y, x, u1);                                   \
            u2 = fma((SGN) * c23.x, x, u2); u3 = fma(-(SGN) * c23.y, x, u3);                                   \
        } else {                                                                                               \
            const double p0 = c01.x * x, p1 = c01.y * x, p2 = c23.x * x, p3 = c23.y * x;                       \
            l0 = l0 + p0; l1 = l1 + p1; l2 = l2 + p2; l3 = l3 + p3;                                            \
            if ((SGN) > 0) { u0 = u0 + p0; u1 = u1 - p1; u2 = u2 + p2; u3 = u3 - p3; }                         \
            else { u0 = u0 - p0; u1 = u1 + p1; u2 = u2 - p2; u3 = u3 + p3; }                                   \
        }                                                                                                      \
    }

#define YR_F2_DUAL_BLOCK(SRC, DL, DU, STRIDE_OUT_L, STRIDE_OUT_U)                                              \
    {                                                                                                          \
        const pair_t* cp = reinterpret_cast<const pair_t*>(Cp + panel_off(b, nt));                             \
        const double* s = (SRC);                                                                               \
        double l0 = 0.0, l1 = 0.0, l2 = 0.0, l3 = 0.0, u0 = 0.0, u1 = 0.0, u2 = 0.0, u3 = 0.0;                 \
        int k = i0;                                                                                            \
        for (; k + 1 < n; k += 2) { YR_F2_DUAL_STEP(1.0) YR_F2_DUAL_STEP(-1.0) }                               \
        if (k < n) YR_F2_DUAL_STEP(1.0)                                                                        \
        const int rem = r - i0;                                                                                \
        double* dl = (DL); double* du = (DU);                                                                  \
        dl[0] = l0; du[0] = u0; asl += fabs(l0); asu += fabs(u0);                                              \
        if (rem > 1) { dl[STRIDE_OUT_L] = l1; du[STRIDE_OUT_U] = u1; asl += fabs(l1); asu += fabs(u1); }       \
        if (rem > 2) { dl[2 * (STRIDE_OUT_L)] = l2; du[2 * (STRIDE_OUT_U)] = u2; asl += fabs(l2); asu += fabs(u2); } \
        if (rem > 3) { dl[3 * (STRIDE_OUT_L)] = l3; du[3 * (STRIDE_OUT_U)] = u3; asl += fabs(l3); asu += fabs(u3); } \
    }

// warp teams: fiber per lane; either output may be src (in place: a block's stores follow all of its loads);
// dstL != dstU; same strides everywhere
template <bool FMA, class Team>
YR_DEV void pass_fiber_dual(Team& tm, const double* src, double* dstL, double* dstU, int sk, int sf, int n, int r, int F,
                            const double* Cp, int nt, unsigned long long& macs, double& sumL, double& sumU) {
    const int B = (r + 3) >> 2;
    double asl = 0.0, asu = 0.0;
    for (int f = tm.tid; f < F; f += tm.nthreads) {
        for (int b = 0; b < B; ++b) {
            const int i0 = b << 2;
            YR_F2_DUAL_BLOCK(src + f * sf + i0 * sk, dstL + f * sf + i0 * sk, dstU + f * sf + i0 * sk, sk, sk)
        }
    }
    if (tm.tid == 0) macs += 2ull * (unsigned)(F * (r * n - r * (r - 1) / 2));
    tm.sum2_sync(asl, asu, sumL, sumU);
}

// CTA teams: items (row block, fiber), out of place
template <bool FMA, class Team>
YR_DEV void restrict_pass_dual(Team& tm, const double* src, int sk, int sf, double* dstL, double* dstU, int dk, int df,
                               int n, int r, int F, const double* Cp, int nt, unsigned long long& macs, double& sumL, double& sumU) {
    const int B = (r + 3) >> 2;
    double asl = 0.0, asu = 0.0;
    const int items = B * F, T = tm.nthreads;
    for (int w0 = 0, wave = 0; w0 < items; w0 += T, ++wave) {                        // snake order, see restrict_pass
        const int id = w0 + ((wave & 1) ? (T - 1 - tm.tid) : tm.tid);
        if (id >= items) continue;
        const int b = id / F, f = id - b * F;
        const int i0 = b << 2;
        YR_F2_DUAL_BLOCK(src + i0 * sk + f * sf, dstL + i0 * dk + f * df, dstU + i0 * dk + f * df, dk, dk)
    }
    if (tm.tid == 0) macs += 2ull * (unsigned)(F * (r * n - r * (r - 1) / 2));
    tm.sum2_sync(asl, asu, sumL, sumU);
}

// ---- RESTRICT ------------------------------------------------------------------------------------------
// slice: [A: T][B: T][X: T][C0 panels][C1 panels][rows_after ints]
// Tensor 0 is fetched into A and tensor 1 into B with asynchronous copies that run behind the matrix build; each
// tensor then ping-pongs with X:  axis 0: own region -> X,  axis 1: X -> own region.
template <bool FMA, class Team>
YR_DEV void restrict_box(Team& tm, const Args& a, unsigned int b, double* slice, unsigned long long& macs) {
    BoxRec<2>& rec = a.recs[b];
    Work& wk = a.work[b];
    const SolverParams& prm = a.prm;
    const int copy_only = wk.copy_only;
    int shp[2][2], ldg[2];
    for (int p = 0; p < 2; ++p) { shp[p][0] = rec.shape[p][0]; shp[p][1] = rec.shape[p][1]; ldg[p] = wk.ld[p]; }
    // shared-memory row stride = the arena's when that is odd (flat copy; a trimmed view keeps its parent's stride)
    const int lds0 = ((ldg[0] & 1) && ldg[0] >= shp[0][1]) ? ldg[0] : odd_ld(shp[0][1]);
    const int lds1 = ((ldg[1] & 1) && ldg[1] >= shp[1][1]) ? ldg[1] : odd_ld(shp[1][1]);
    const int sz0 = tensor_doubles(shp[0][0], lds0), sz1 = tensor_doubles(shp[1][0], lds1);
    const int T = sz0 > sz1 ? sz0 : sz1;
    const int n0max = shp[0][0] > shp[1][0] ? shp[0][0] : shp[1][0];
    const int n1max = shp[0][1] > shp[1][1] ? shp[0][1] : shp[1][1];
    double* A = slice; double* Bf = slice + T; double* X = slice + 2 * T;
    double* C0 = slice + 3 * T; double* C1 = C0 + panel_doubles(n0max);
    int* rows0 = reinterpret_cast<int*>(C1 + panel_doubles(n1max)); int* rows1 = rows0 + n0max;
    const double al0 = wk.alpha[0], be0 = wk.beta[0], al1 = wk.alpha[1], be1 = wk.beta[1];
    const bool id0 = copy_only || (al0 == 1.0 && be0 == 0.0), id1 = copy_only || (al1 == 1.0 && be1 == 0.0);
    const unsigned long long off_in[2] = {wk.off[0], wk.off[1]};
    double sum_in[2] = {wk.sumabs[0], wk.sumabs[1]};
    double err_in[2] = {rec.err[0], rec.err[1]};
    // output space: never larger than the input (rows only shrink)
    unsigned long long base = 0;
    if (tm.tid == 0) base = tm.atomic_add(&a.ctl->data_used, (unsigned long long)(sz0 + sz1));
    base = tm.bcast(base);
    if (base + (unsigned long long)(sz0 + sz1) > a.cap_data_out) { if (tm.tid == 0) tm.atomic_add(&a.ctl->overflow, 1ull); return; }
    tm.sync();                                                   // the previous box is done with the slice
    // fetch both tensors (A <- tensor 0, B <- tensor 1) behind the matrix build
    for (int p = 0; p < 2; ++p) {
        double* dst = p ? Bf : A;
        const int lds = p ? lds1 : lds0, n0 = shp[p][0], n1 = shp[p][1];
        const double* src = a.data_in + off_in[p];
        if (ldg[p] == lds && (off_in[p] & 1ull) == 0) tm.copy_async(dst, src, tensor_doubles(n0, lds));
        else for (int e = tm.tid; e < n0 * n1; e += tm.nthreads) { const int i = e / n1, j = e - i * n1; dst[i * lds + j] = src[(long long)i * ldg[p] + j]; }
    }
    {
        PanelJob j0 = {C0, rows0, (!id0 && n0max > 1) ? n0max : 0, al0, be0};
        PanelJob j1 = {C1, rows1, (!id1 && n1max > 1) ? n1max : 0, al1, be1};
        if (j0.nt > 0 || j1.nt > 0) build_two(tm, j0, j1);
    }
    tm.copy_wait();
    tm.sync();
    unsigned long long out_off = base;
    for (int p = 0; p < 2; ++p) {
        const int n0 = shp[p][0], n1 = shp[p][1];
        const int lds = p ? lds1 : lds0;
        double* cur = p ? Bf : A; double* oth = X;
        double s = copy_only ? view_abs_sum(tm, cur, n0, n1, lds) : sum_in[p];
        double err = err_in[p];
        int r0 = n0, r1 = n1, ld = lds;
        if (!copy_only) err += n0 * kMachEps * s;                    // charged before the axis (:860)
        if (!id0 && n0 > 1) {
            r0 = rows0[n0 - 1];
            s = restrict_pass<FMA>(tm, cur, ld, 1, oth, ld, 1, n0, r0, n1, C0, n0max, macs);
            double* t = cur; cur = oth; oth = t;
        }
        if (!copy_only) err += n1 * kMachEps * s;
        if (!id1 && n1 > 1) {
            r1 = rows1[n1 - 1];
            const int ld2 = odd_ld(r1);
            s = restrict_pass<FMA>(tm, cur, 1, ld, oth, 1, ld2, n1, r1, r0, C1, n1max, macs);
            double* t = cur; cur = oth; oth = t;
            ld = ld2;
        }
        const int osz = tensor_doubles(r0, ld);
        if (tm.warp == 0) {
            Epilogue ep;
            tensor_epilogue(tm, cur, r0, r1, ld, s, err, prm, ep);
            if (tm.tid == 0) {
                store_epilogue(wk, p, ep, s);
                rec.shape[p][0] = r0; rec.shape[p][1] = r1; rec.err[p] = err;
                wk.off[p] = out_off; wk.ld[p] = ld;
            }
        }
        copy_flat(tm, a.data_out + out_off, cur, osz);
        out_off += (unsigned long long)osz;
        tm.sync();                                               // tensor 0 has left X before tensor 1's passes overwrite it
    }
    if (tm.tid == 0) { wk.op = kTNone; wk.copy_only = 0; wk.trim_valid = 1; }
}

// ---- SUBDIVIDE -----------------------------------------------------------------------------------------
// slice: [S][L][U][X], each T doubles.  Matrices come from the level-independent tables (shared-memory copy
// `smat` for m = 0 when the launch staged one, else the global table).
struct SubMats { const double* lo; const double* hi; const int* rlo; const int* rhi; int nt; };

template <class Team>
YR_DEV void child_record(Team& tm, const Args& a, const BoxRec<2>& rec, const Work& wk, unsigned int child, int c, int k,
                         const int split_axes[2]) {
    // :1111-1128 -- thread-serial
    BoxRec<2>& out = a.recs[child];
    Work& ow = a.work[child];
    // every array index below is a compile-time constant (selects instead of indexing by the axis), so the interval
    // state stays in registers: indexed by a run-time axis it lives on the stack, one local-memory trip per access
    IntervalState<2> st = rec.st;
    const int ax0 = split_axes[0], ax1 = split_axes[1];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        if (j < k) {
            const int ax = j ? ax1 : ax0;
            const int half = (c >> (k - 1 - j)) & 1;
            const double mid = ax ? st.next_mid[1] : st.next_mid[0];
            double sub[2][2], al[2], be[2];
            sub[0][0] = (ax == 0 && half) ? mid : -1.0; sub[0][1] = (ax == 0 && !half) ? mid : 1.0;
            sub[1][0] = (ax == 1 && half) ? mid : -1.0; sub[1][1] = (ax == 1 && !half) ? mid : 1.0;
            add_transform<2>(st, sub, al, be);
        }
    }
    if (k > 0) { if (ax0 == 0) st.next_mid[0] = 0.0; else st.next_mid[1] = 0.0; }
    if (k > 1) { if (ax1 == 0) st.next_mid[0] = 0.0; else st.next_mid[1] = 0.0; }
    out.st = st;
    out.data_off = 0;
    out.system = rec.system;
    const bool fsplit = wk.final_split != 0;
    out.parent_node = fsplit ? rec.parent_node : wk.node_index;
    out.collector = fsplit ? wk.result_index : rec.collector;
    out.level = wk.node_level;
    unsigned long long hi = rec.path_hi, lo = rec.path_lo;
    const int used = rec.path_bits;
    const int shift = 128 - used - k;
    if (shift >= 64) hi |= (unsigned long long)c << (shift - 64);
    else if (shift >= 0) {
        lo |= (unsigned long long)c << shift;
        if (shift + k > 64) hi |= (unsigned long long)c >> (64 - shift);
    }
    out.path_bits = (shift >= 0) ? used + k : used;
    out.pad0 = 0;
    out.path_hi = hi; out.path_lo = lo;
    for (int d = 0; d < 2; ++d) { ow.cell_iv[d][0] = st.iv[d][0]; ow.cell_iv[d][1] = st.iv[d][1]; }
    ow.op = kTNone; ow.fresh = 1; ow.copy_only = 0; ow.trim_valid = 1;
    ow.zoom_count = 0; ow.changed = 1; ow.should_stop = 0; ow.final_split = 0;
    ow.node_level = wk.node_level; ow.result_index = -1; ow.node_index = -1; ow.nsplit = 0;
    (void)tm;
}

template <bool FMA, class Team>
YR_DEV void subdivide_box(Team& tm, const Args& a, unsigned int b, double* slice, const double* smat_lo, const double* smat_hi,
                          int smat_nt, unsigned long long& macs) {
    const BoxRec<2>& rec = a.recs[b];
    const Work& wk = a.work[b];
    const SolverParams& prm = a.prm;
    const int k = wk.nsplit, nchild = 1 << k;
    int shp[2][2];
    for (int p = 0; p < 2; ++p) { shp[p][0] = rec.shape[p][0]; shp[p][1] = rec.shape[p][1]; }
    const int sz0 = tensor_doubles(shp[0][0], wk.ld[0]), sz1 = tensor_doubles(shp[1][0], wk.ld[1]);
    const int T = sz0 > sz1 ? sz0 : sz1;
    int split_axes[2]; { bool used[2] = {false, false}; for (int j = 0; j < k; ++j) used[wk.order[0][j]] = true; int j = 0; for (int ax = 0; ax < 2; ++ax) if (used[ax]) split_axes[j++] = ax; if (j < 2) split_axes[1] = split_axes[0]; }
    // which table: m = 0 (every split after an axis' first) or m = kFirstSplit
    int midsel[2];
    for (int ax = 0; ax < 2; ++ax) midsel[ax] = (rec.st.next_mid[ax] == 0.0) ? 0 : 1;
    // slots and output space: children are never larger than their parent
    unsigned long long cbase = 0, dbase = 0;
    const unsigned long long per_child = (unsigned long long)(sz0 + sz1);
    if (tm.tid == 0) {
        cbase = tm.atomic_add(&a.ctl->n_boxes, (unsigned long long)nchild);
        dbase = tm.atomic_add(&a.ctl->data_used, per_child * nchild);
        for (int j = 0; j < k; ++j) { const double m = rec.st.next_mid[split_axes[j]]; if (m != 0.0 && m != kFirstSplit) tm.atomic_add(&a.ctl->bad_mid, 1ull); }
    }
    cbase = tm.bcast(cbase); dbase = tm.bcast(dbase);
    if (cbase + nchild > a.cap_boxes || dbase + per_child * nchild > a.cap_data_out) { if (tm.tid == 0) tm.atomic_add(&a.ctl->overflow, 1ull); return; }
    if (tm.tid < nchild) child_record(tm, a, rec, wk, (unsigned int)(cbase + tm.tid), tm.tid, k, split_axes);
    double* S = slice; double* L = slice + T; double* U = slice + 2 * T; double* X = slice + 3 * T;
    for (int p = 0; p < 2; ++p) {
        const int n0 = shp[p][0], n1 = shp[p][1], ld = wk.ld[p];
        tm.sync();
        if ((wk.off[p] & 1ull) == 0) copy_flat(tm, S, a.data_in + wk.off[p], tensor_doubles(n0, ld));
        else for (int e = tm.tid; e < n0 * ld; e += tm.nthreads) S[e] = a.data_in[wk.off[p] + e];
        tm.sync();
        // stage 1 along a1 = order[p][0]: S -> L (lower half), U (upper half)
        const int a1 = wk.order[p][0];
        const int n_a1 = a1 == 0 ? n0 : n1;
        SubMats m1;
        {
            const int ms = midsel[a1];
            const bool sm = (ms == 0 && smat_lo != nullptr);
            m1.lo = sm ? smat_lo : a.tab.mat[ms][0]; m1.hi = sm ? smat_hi : a.tab.mat[ms][1];
            m1.rlo = a.tab.rows_after[ms][0]; m1.rhi = a.tab.rows_after[ms][1]; m1.nt = sm ? smat_nt : a.tab.nt;
        }
        const double e1 = n_a1 * kMachEps * wk.sumabs[p];
        const double terr1 = e1 + rec.err[p];
        int shL[2] = {n0, n1}, shU[2] = {n0, n1}; int ldL = ld, ldU = ld;
        double sumL, sumU;
        if (n_a1 > 1) {
            const int rl = m1.rlo[n_a1 - 1], ru = m1.rhi[n_a1 - 1];
            const bool mirror = (midsel[a1] == 0) && (rl == ru);      // m = 0: C_hi = +-C_lo, one sweep for both
            if (a1 == 0) {
                shL[0] = rl; shU[0] = ru;
                if (mirror) restrict_pass_dual<FMA>(tm, S, ld, 1, L, U, ld, 1, n0, rl, n1, m1.lo, m1.nt, macs, sumL, sumU);
                else {
                    sumL = restrict_pass<FMA>(tm, S, ld, 1, L, ld, 1, n0, rl, n1, m1.lo, m1.nt, macs);
                    sumU = restrict_pass<FMA>(tm, S, ld, 1, U, ld, 1, n0, ru, n1, m1.hi, m1.nt, macs);
                }
            } else {
                shL[1] = rl; shU[1] = ru; ldL = odd_ld(rl); ldU = odd_ld(ru);
                if (mirror) restrict_pass_dual<FMA>(tm, S, 1, ld, L, U, 1, ldL, n1, rl, n0, m1.lo, m1.nt, macs, sumL, sumU);
                else {
                    sumL = restrict_pass<FMA>(tm, S, 1, ld, L, 1, ldL, n1, rl, n0, m1.lo, m1.nt, macs);
                    sumU = restrict_pass<FMA>(tm, S, 1, ld, U, 1, ldU, n1, ru, n0, m1.hi, m1.nt, macs);
                }
            }
        } else {
            // extent-1 axis: TransformChebInPlaceND returns the tensor unchanged (:363)
            copy_flat(tm, L, S, tensor_doubles(n0, ld)); copy_flat(tm, U, S, tensor_doubles(n0, ld));
            sumL = sumU = wk.sumabs[p];
            tm.sync();
        }
        const int rank0 = (split_axes[0] == a1) ? 0 : 1;      // position of a1 in the ascending set of split axes
        for (int h1 = 0; h1 < 2; ++h1) {
            const double* H = h1 ? U : L;
            const int* shH = h1 ? shU : shL; const int ldH = h1 ? ldU : ldL; const double sumH = h1 ? sumU : sumL;
            if (k == 1) {
                const int c = h1;
                const int osz = tensor_doubles(shH[0], ldH);
                const unsigned long long off = dbase + per_child * c + (p ? (unsigned long long)sz0 : 0ull);
                if (tm.warp == 0) {
                    Epilogue ep;
                    tensor_epilogue(tm, H, shH[0], shH[1], ldH, sumH, terr1, prm, ep);
                    if (tm.tid == 0) {
                        const unsigned int child = (unsigned int)(cbase + c);
                        Work& ow = a.work[child]; BoxRec<2>& orec = a.recs[child];
                        store_epilogue(ow, p, ep, sumH);
                        orec.shape[p][0] = shH[0]; orec.shape[p][1] = shH[1]; orec.err[p] = terr1;
                        ow.off[p] = off; ow.ld[p] = ldH;
                    }
                }
                copy_flat(tm, a.data_out + off, H, osz);
                continue;
            }
            // stage 2 along a2 = order[p][1]: H -> (lower -> S, upper -> X)
            const int a2 = wk.order[p][1];
            const int n_a2 = shH[a2];
            SubMats m2;
            {
                const int ms = midsel[a2];
                const bool sm = (ms == 0 && smat_lo != nullptr);
                m2.lo = sm ? smat_lo : a.tab.mat[ms][0]; m2.hi = sm ? smat_hi : a.tab.mat[ms][1];
                m2.rlo = a.tab.rows_after[ms][0]; m2.rhi = a.tab.rows_after[ms][1]; m2.nt = sm ? smat_nt : a.tab.nt;
            }
            const double e2 = n_a2 * kMachEps * sumH;
            const double terr2 = e2 + terr1;
            int shA[2] = {shH[0], shH[1]}, shB[2] = {shH[0], shH[1]}; int ldA = ldH, ldB = ldH;
            double sumA, sumB;
            tm.sync();                                         // S / X are free (previous epilogue copies are done)
            if (n_a2 > 1) {
                const int rl = m2.rlo[n_a2 - 1], ru = m2.rhi[n_a2 - 1];
                const bool mirror = (midsel[a2] == 0) && (rl == ru);
                if (a2 == 0) {
                    shA[0] = rl; shB[0] = ru;
                    if (mirror) restrict_pass_dual<FMA>(tm, H, ldH, 1, S, X, ldH, 1, shH[0], rl, shH[1], m2.lo, m2.nt, macs, sumA, sumB);
                    else {
                        sumA = restrict_pass<FMA>(tm, H, ldH, 1, S, ldH, 1, shH[0], rl, shH[1], m2.lo, m2.nt, macs);
                        sumB = restrict_pass<FMA>(tm, H, ldH, 1, X, ldH, 1, shH[0], ru, shH[1], m2.hi, m2.nt, macs);
                    }
                } else {
                    shA[1] = rl; shB[1] = ru; ldA = odd_ld(rl); ldB = odd_ld(ru);
                    if (mirror) restrict_pass_dual<FMA>(tm, H, 1, ldH, S, X, 1, ldA, shH[1], rl, shH[0], m2.lo, m2.nt, macs, sumA, sumB);
                    else {
                        sumA = restrict_pass<FMA>(tm, H, 1, ldH, S, 1, ldA, shH[1], rl, shH[0], m2.lo, m2.nt, macs);
                        sumB = restrict_pass<FMA>(tm, H, 1, ldH, X, 1, ldB, shH[1], ru, shH[0], m2.hi, m2.nt, macs);
                    }
                }
            } else {
                copy_flat(tm, S, H, tensor_doubles(shH[0], ldH)); copy_flat(tm, X, H, tensor_doubles(shH[0], ldH));
                sumA = sumB = sumH;
                tm.sync();
            }
            for (int h2 = 0; h2 < 2; ++h2) {
                const double* O = h2 ? X : S;
                const int* shO = h2 ? shB : shA; const int ldO = h2 ? ldB : ldA; const double sumO = h2 ? sumB : sumA;
                // canonical child: split axes ascending, smallest axis most significant
                const int c = (rank0 == 0) ? ((h1 << 1) | h2) : ((h2 << 1) | h1);
                const int osz = tensor_doubles(shO[0], ldO);
                const unsigned long long off = dbase + per_child * c + (p ? (unsigned long long)sz0 : 0ull);
                if (tm.warp == 0) {
                    Epilogue ep;
                    tensor_epilogue(tm, O, shO[0], shO[1], ldO, sumO, terr2, prm, ep);
                    if (tm.tid == 0) {
                        const unsigned int child = (unsigned int)(cbase + c);
                        Work& ow = a.work[child]; BoxRec<2>& orec = a.recs[child];
                        store_epilogue(ow, p, ep, sumO);
                        orec.shape[p][0] = shO[0]; orec.shape[p][1] = shO[1]; orec.err[p] = terr2;
                        ow.off[p] = off; ow.ld[p] = ldO;
                    }
                }
                copy_flat(tm, a.data_out + off, O, osz);
            }
        }
    }
}

// ---- warp teams: in-place, fiber per lane -----------------------------------------------------------------
// Lane f owns fiber f and walks its row blocks in ascending order: block b reads source rows >= 4b and then
// overwrites rows 4b..4b+3, which no later block reads -- so a pass can run in place (dst == src), every lane
// has the same trip count, and a box needs one tensor region instead of three.  dst may also be another region
// (same strides).  Returns the team-wide sum|dst|; ends with the team barrier.
template <bool FMA, class Team>
YR_DEV double pass_fiber(Team& tm, const double* src, double* dst, int sk, int sf, int n, int r, int F,
                         const double* Cp, int nt, unsigned long long& macs) {
    const int B = (r + 3) >> 2;
    double asum = 0.0;
    for (int f = tm.tid; f < F; f += tm.nthreads) {
        const double* sc = src + f * sf;
        double* dc = dst + f * sf;
        for (int b = 0; b < B; ++b) {
            const int i0 = b << 2;
            const pair_t* cp = reinterpret_cast<const pair_t*>(Cp + panel_off(b, nt));
            const double* s = sc + i0 * sk;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 2
            for (int k = i0; k < n; ++k) {
                const double x = *s; s += sk;
                const pair_t c01 = cp[0], c23 = cp[1]; cp += 2;
                if (FMA) { a0 = fma(c01.x, x, a0); a1 = fma(c01.y, x, a1); a2 = fma(c23.x, x, a2); a3 = fma(c23.y, x, a3); }
                else { a0 = a0 + c01.x * x; a1 = a1 + c01.y * x; a2 = a2 + c23.x * x; a3 = a3 + c23.y * x; }
            }
            double* d = dc + i0 * sk;
            const int rem = r - i0;
            d[0] = a0; asum += fabs(a0);
            if (rem > 1) { d[sk] = a1; asum += fabs(a1); }
            if (rem > 2) { d[2 * sk] = a2; asum += fabs(a2); }
            if (rem > 3) { d[3 * sk] = a3; asum += fabs(a3); }
        }
    }
    if (tm.tid == 0) macs += (unsigned long long)(unsigned)(F * (r * n - r * (r - 1) / 2));
    return tm.sum_sync(asum);
}

// tensor view (r0 x r1, row stride ld) in shared memory -> arena, re-laid out with row stride ld_out
template <class Team>
YR_DEV void copy_out_rows(Team& tm, double* dst, int ld_out, const double* M, int ld, int r0, int r1) {
    if (ld_out == ld) { tm.store_async(dst, M, tensor_doubles(r0, ld)); return; }      // asynchronous: tm.store_wait() before M is reused
    // four lanes per row (one 32-byte sector per store), nthreads / 4 rows per sweep: two counted loops, no index
    // arithmetic per element
    const int cg = tm.nthreads >= 4 ? 4 : tm.nthreads;             // lanes per row (teams are powers of two)
    const int rows_per_sweep = tm.nthreads / cg, a = tm.tid / cg, b = tm.tid - a * cg;
    for (int i = a; i < r0; i += rows_per_sweep) {
        const double* s = M + i * ld;
        double* d = dst + i * ld_out;
        for (int j = b; j < r1; j += cg) d[j] = s[j];
    }
}

template <class Team>
YR_DEVN void fetch_strided(Team& tm, double* dst, int lds, const double* src, int ldg, int n0, int n1) {   // level-0 boxes only
    for (int i = 0; i < n0; ++i) for (int j = tm.tid; j < n1; j += tm.nthreads) dst[i * lds + j] = src[(long long)i * ldg + j];
}
template <class Team>
YR_DEV void fetch_tensor(Team& tm, double* dst, int lds, const double* src, int ldg, int n0, int n1, bool flat) {
    if (flat) tm.copy_async(dst, src, tensor_doubles(n0, lds));
    else fetch_strided(tm, dst, lds, src, ldg, n0, n1);
}

// RESTRICT, warp team.  slice: [A: T][C0 panels][C1 panels][rows_after ints]
template <bool FMA, class Team>
YR_DEV void restrict_box_w(Team& tm, const Args& a, unsigned int b, double* slice, unsigned long long& macs) {
    BoxRec<2>& rec = a.recs[b];
    Work& wk = a.work[b];
    const SolverParams& prm = a.prm;
    const int copy_only = wk.copy_only;
    // per-tensor geometry in scalars, selected by p below: as arrays indexed by the loop variable of the (not unrolled)
    // tensor loop they were kept in local memory (a 144-byte stack frame in the 16-lane RESTRICT kernel)
    const int s00 = rec.shape[0][0], s01 = rec.shape[0][1], s10 = rec.shape[1][0], s11 = rec.shape[1][1];
    const int ldg0 = wk.ld[0], ldg1 = wk.ld[1];
    const unsigned long long off0 = wk.off[0], off1 = wk.off[1];
    const bool flat0 = (ldg0 & 1) && ldg0 >= s01 && (off0 & 1ull) == 0, flat1 = (ldg1 & 1) && ldg1 >= s11 && (off1 & 1ull) == 0;
    const int lds0 = flat0 ? ldg0 : odd_ld(s01), lds1 = flat1 ? ldg1 : odd_ld(s11);
    const int sz0 = tensor_doubles(s00, lds0), sz1 = tensor_doubles(s10, lds1);
    const int T = sz0 > sz1 ? sz0 : sz1;
    const int n0max = s00 > s10 ? s00 : s10;
    const int n1max = s01 > s11 ? s01 : s11;
    double* A = slice;
    double* C0 = slice + T; double* C1 = C0 + panel_doubles(n0max);
    int* rows0 = reinterpret_cast<int*>(C1 + panel_doubles(n1max)); int* rows1 = rows0 + n0max;
    const double al0 = wk.alpha[0], be0 = wk.beta[0], al1 = wk.alpha[1], be1 = wk.beta[1];
    const bool id0 = copy_only || (al0 == 1.0 && be0 == 0.0), id1 = copy_only || (al1 == 1.0 && be1 == 0.0);
    const double sum_in0 = wk.sumabs[0], sum_in1 = wk.sumabs[1];
    const double err_in0 = rec.err[0], err_in1 = rec.err[1];
    unsigned long long base = 0;
    if (tm.tid == 0) base = tm.atomic_add(&a.ctl->data_used, (unsigned long long)(sz0 + sz1));
    base = tm.bcast(base);
    if (base + (unsigned long long)(sz0 + sz1) > a.cap_data_out) { if (tm.tid == 0) tm.atomic_add(&a.ctl->overflow, 1ull); return; }
    tm.sync();                                                   // the previous box is done with the slice
    fetch_tensor(tm, A, lds0, a.data_in + off0, ldg0, s00, s01, flat0);   // runs behind the build
    {
        PanelJob j0 = {C0, rows0, (!id0 && n0max > 1) ? n0max : 0, al0, be0};
        PanelJob j1 = {C1, rows1, (!id1 && n1max > 1) ? n1max : 0, al1, be1};
        if (j0.nt > 0 || j1.nt > 0) build_two(tm, j0, j1);
    }
    unsigned long long out_off = base;
    for (int p = 0; p < 2; ++p) {
        const int n0 = p ? s10 : s00, n1 = p ? s11 : s01, ld = p ? lds1 : lds0;
        if (p == 1) fetch_tensor(tm, A, ld, a.data_in + off1, ldg1, n0, n1, flat1);
        tm.copy_wait();
        tm.sync();
        double s = copy_only ? view_abs_sum(tm, A, n0, n1, ld) : (p ? sum_in1 : sum_in0);
        double err = p ? err_in1 : err_in0;
        int r0 = n0, r1 = n1;
#pragma unroll 1
        for (int axis = 0; axis < 2; ++axis) {                       // one copy of the pass in the instruction stream
            const int n = axis ? n1 : n0;
            if (!copy_only) err += n * kMachEps * s;                 // charged before the axis (:860)
            if ((axis ? id1 : id0) || n <= 1) continue;
            const int r = (axis ? rows1 : rows0)[n - 1];
            s = pass_fiber<FMA>(tm, A, A, axis ? 1 : ld, axis ? ld : 1, n, r, axis ? r0 : n1, axis ? C1 : C0, axis ? n1max : n0max, macs);
            if (axis) r1 = r; else r0 = r;
        }
        // re-laying the tensor out row by row only pays when it saves real space: keep the stride for a small gap
        const int ld_out = (ld - odd_ld(r1) <= 2) ? ld : odd_ld(r1), osz = tensor_doubles(r0, ld_out);
        Epilogue ep;
        tensor_epilogue(tm, A, r0, r1, ld, s, err, prm, ep);
        copy_out_rows(tm, a.data_out + out_off, ld_out, A, ld, r0, r1);
        if (tm.tid == 0) {
            store_epilogue(wk, p, ep, s);
            rec.shape[p][0] = r0; rec.shape[p][1] = r1; rec.err[p] = err;
            wk.off[p] = out_off; wk.ld[p] = ld_out;
        }
        out_off += (unsigned long long)osz;
        tm.store_wait();
        tm.sync();                                               // tensor 0 has left A before tensor 1 arrives
    }
    if (tm.tid == 0) { wk.op = kTNone; wk.copy_only = 0; wk.trim_valid = 1; }
}

// SUBDIVIDE, warp team.  slice: [S][L][X], each T doubles; every region keeps the source's row stride.
//   stage 1:  L = lower(S) out of place, then S = upper(S) in place
//   stage 2 (from H = L, then S):  X = upper(H) out of place, then H = lower(H) in place; both leave for the arena
template <bool FMA, class Team>
YR_DEV void subdivide_box_w(Team& tm, const Args& a, unsigned int b, double* slice, const double* smat_lo, const double* smat_hi,
                            int smat_nt, unsigned long long& macs) {
    const BoxRec<2>& rec = a.recs[b];
    const Work& wk = a.work[b];
    const SolverParams& prm = a.prm;
    const int k = wk.nsplit, nchild = 1 << k;
    int shp[2][2];
    for (int p = 0; p < 2; ++p) { shp[p][0] = rec.shape[p][0]; shp[p][1] = rec.shape[p][1]; }
    const int sz0 = tensor_doubles(shp[0][0], wk.ld[0]), sz1 = tensor_doubles(shp[1][0], wk.ld[1]);
    const int T = sz0 > sz1 ? sz0 : sz1;
    int split_axes[2]; { bool used[2] = {false, false}; for (int j = 0; j < k; ++j) used[wk.order[0][j]] = true; int j = 0; for (int ax = 0; ax < 2; ++ax) if (used[ax]) split_axes[j++] = ax; if (j < 2) split_axes[1] = split_axes[0]; }
    int midsel[2];
    for (int ax = 0; ax < 2; ++ax) midsel[ax] = (rec.st.next_mid[ax] == 0.0) ? 0 : 1;
    unsigned long long cbase = 0, dbase = 0;
    const unsigned long long per_child = (unsigned long long)(sz0 + sz1);
    if (tm.tid == 0) {
        cbase = tm.atomic_add(&a.ctl->n_boxes, (unsigned long long)nchild);
        dbase = tm.atomic_add(&a.ctl->data_used, per_child * nchild);
        for (int j = 0; j < k; ++j) { const double m = rec.st.next_mid[split_axes[j]]; if (m != 0.0 && m != kFirstSplit) tm.atomic_add(&a.ctl->bad_mid, 1ull); }
    }
    cbase = tm.bcast(cbase); dbase = tm.bcast(dbase);
    if (cbase + nchild > a.cap_boxes || dbase + per_child * nchild > a.cap_data_out) { if (tm.tid == 0) tm.atomic_add(&a.ctl->overflow, 1ull); return; }
    tm.sync();
    tm.copy_async(slice, a.data_in + wk.off[0], sz0);           // tensor 0 arrives while the child records are written
    if (tm.tid < nchild) child_record(tm, a, rec, wk, (unsigned int)(cbase + tm.tid), tm.tid, k, split_axes);
    double* S = slice; double* L = slice + T; double* X = slice + 2 * T;
    unsigned dead_children = 0;                                  // children already known to fail the constant check
    for (int p = 0; p < 2; ++p) {
        const int n0 = shp[p][0], n1 = shp[p][1], ld = wk.ld[p];
        if (p == 1) tm.copy_async(S, a.data_in + wk.off[1], sz1);
        tm.copy_wait();
        tm.sync();
        const int a1 = wk.order[p][0];
        const int n_a1 = a1 == 0 ? n0 : n1;
        SubMats m1;
        {
            const int ms = midsel[a1];
            const bool sm = (ms == 0 && smat_lo != nullptr);
            m1.lo = sm ? smat_lo : a.tab.mat[ms][0]; m1.hi = sm ? smat_hi : a.tab.mat[ms][1];
            m1.rlo = a.tab.rows_after[ms][0]; m1.rhi = a.tab.rows_after[ms][1]; m1.nt = sm ? smat_nt : a.tab.nt;
        }
        const double e1 = n_a1 * kMachEps * wk.sumabs[p];
        const double terr1 = e1 + rec.err[p];
        int shL[2] = {n0, n1}, shU[2] = {n0, n1};
        double sumL, sumU;
        if (n_a1 > 1) {
            const int rl = m1.rlo[n_a1 - 1], ru = m1.rhi[n_a1 - 1];
            shL[a1] = rl; shU[a1] = ru;
            const int sk = a1 == 0 ? ld : 1, sf = a1 == 0 ? 1 : ld, F = a1 == 0 ? n1 : n0;
            if (midsel[a1] == 0 && rl == ru) pass_fiber_dual<FMA>(tm, S, L, S, sk, sf, n_a1, rl, F, m1.lo, m1.nt, macs, sumL, sumU);
            else {
                sumL = pass_fiber<FMA>(tm, S, L, sk, sf, n_a1, rl, F, m1.lo, m1.nt, macs);
                sumU = pass_fiber<FMA>(tm, S, S, sk, sf, n_a1, ru, F, m1.hi, m1.nt, macs);
            }
        } else {
            copy_flat(tm, L, S, tensor_doubles(n0, ld));         // extent-1 axis: unchanged (:363)
            sumL = sumU = wk.sumabs[p];
            tm.sync();
        }
        const int rank0 = (split_axes[0] == a1) ? 0 : 1;
        for (int h1 = 0; h1 < 2; ++h1) {
            double* H = h1 ? S : L;
            const int* shH = h1 ? shU : shL; const double sumH = h1 ? sumU : sumL;
            double* outs[2] = {H, X}; int shO[2][2] = {{shH[0], shH[1]}, {shH[0], shH[1]}}; double sumO[2] = {sumH, sumH};
            double terr = terr1;
            int nout = 1;
            if (k == 2) {
                const int a2 = wk.order[p][1];
                const int n_a2 = shH[a2];
                SubMats m2;
                {
                    const int ms = midsel[a2];
                    const bool sm = (ms == 0 && smat_lo != nullptr);
                    m2.lo = sm ? smat_lo : a.tab.mat[ms][0]; m2.hi = sm ? smat_hi : a.tab.mat[ms][1];
                    m2.rlo = a.tab.rows_after[ms][0]; m2.rhi = a.tab.rows_after[ms][1]; m2.nt = sm ? smat_nt : a.tab.nt;
                }
                const double e2 = n_a2 * kMachEps * sumH;
                terr = e2 + terr1;
                nout = 2;
                if (n_a2 > 1) {
                    const int rl = m2.rlo[n_a2 - 1], ru = m2.rhi[n_a2 - 1];
                    shO[0][a2] = rl; shO[1][a2] = ru;
                    const int sk = a2 == 0 ? ld : 1, sf = a2 == 0 ? 1 : ld, F = a2 == 0 ? shH[1] : shH[0];
                    // lower stays in H (in place), upper goes to X
                    if (midsel[a2] == 0 && rl == ru) pass_fiber_dual<FMA>(tm, H, H, X, sk, sf, n_a2, rl, F, m2.lo, m2.nt, macs, sumO[0], sumO[1]);
                    else {
                        sumO[1] = pass_fiber<FMA>(tm, H, X, sk, sf, n_a2, ru, F, m2.hi, m2.nt, macs);
                        sumO[0] = pass_fiber<FMA>(tm, H, H, sk, sf, n_a2, rl, F, m2.lo, m2.nt, macs);
                    }
                } else {
                    copy_flat(tm, X, H, tensor_doubles(shH[0], ld));
                    tm.sync();
                }
            }
            for (int h2 = 0; h2 < nout; ++h2) {
                // canonical child: split axes ascending, smallest axis most significant
                const int c = (k == 1) ? h1 : ((rank0 == 0) ? ((h1 << 1) | h2) : ((h2 << 1) | h1));
                const int r0 = shO[h2][0], r1 = shO[h2][1];
                const int ld_out = odd_ld(r1);        // children are re-laid out tightly (keeping the parent's stride: +3 % step time, profiles/r02_experiments.txt)
                const unsigned long long off = dbase + per_child * c + (p ? (unsigned long long)sz0 : 0ull);
                Epilogue ep;
                // A child the constant check (:1213) is going to discard needs neither its trim outcome nor its tensors
                // in the arena: the scalar step repeats the same comparison on the same numbers (sum|M|, c0, error)
                // and drops the box before anything else is read.  About a quarter of all children end here.
                const double c0 = outs[h2][0];
                const bool dead_now = YR_F2_SKIP_DEAD && prm.constant_check && (fabs(c0) > sumO[h2] - fabs(c0) + terr);
                if (dead_now) dead_children |= 1u << c;
                const bool dead = (dead_children >> c) & 1u;
                if (!dead) {
                    tensor_epilogue(tm, outs[h2], r0, r1, ld, sumO[h2], terr, prm, ep);
                    copy_out_rows(tm, a.data_out + off, ld_out, outs[h2], ld, r0, r1);
                } else {
                    for (int q = 0; q < 6; ++q) ep.low[q] = 0.0;
                    ep.low[0] = c0; ep.t0 = r0; ep.t1 = r1; ep.tsum = sumO[h2]; ep.terr = terr;
                }
                if (tm.tid == 0) {
                    const unsigned int child = (unsigned int)(cbase + c);
                    Work& ow = a.work[child]; BoxRec<2>& orec = a.recs[child];
                    store_epilogue(ow, p, ep, sumO[h2]);
                    orec.shape[p][0] = r0; orec.shape[p][1] = r1; orec.err[p] = terr;
                    ow.off[p] = off; ow.ld[p] = ld_out;
                }
            }
            tm.store_wait();
            tm.sync();                                           // X (and H) have left for the arena
        }
    }
}

// ---- scalar step (thread per box) -------------------------------------------------------------------
YR_HD void box_sizes(const BoxRec<2>& rec, const Work& wk, int& T, int& n0max, int& n1max, unsigned long long& total) {
    const int l0 = odd_ld(rec.shape[0][1]) > wk.ld[0] ? odd_ld(rec.shape[0][1]) : wk.ld[0];
    const int l1 = odd_ld(rec.shape[1][1]) > wk.ld[1] ? odd_ld(rec.shape[1][1]) : wk.ld[1];
    const int s0 = tensor_doubles(rec.shape[0][0], l0), s1 = tensor_doubles(rec.shape[1][0], l1);
    T = s0 > s1 ? s0 : s1; total = (unsigned long long)s0 + s1;
    n0max = rec.shape[0][0] > rec.shape[1][0] ? rec.shape[0][0] : rec.shape[1][0];
    n1max = rec.shape[0][1] > rec.shape[1][1] ? rec.shape[0][1] : rec.shape[1][1];
}

struct Decision {
    int op, cls;                                // next tensor op (kTNone: the box is finished)
    unsigned long long need, bound, ext;        // its shared-memory need, output bound, largest extent
    unsigned long long boxes, zooms, dconst, dquad, dzoom, coeffs, maxlevel, subdiv;
};

// `rec` / `wk` are the box's records -- in place (emulator) or a thread-private copy that the caller loads in one
// burst and writes back only when the box lives on (CUDA: the step is a long chain of small dependent accesses,
// which on scattered global records is one memory round trip after another).
template <class At>
YR_HDM void scalar_step_on(const Args& a, BoxRec<2>& rec, Work& wk, Decision& dc, At at) {
    const SolverParams& prm = a.prm;
    IntervalState<2>& st = rec.st;

    auto emit = [&](uint32_t extra_flags) {
        const unsigned long long idx = at.add(&a.ctl->n_results, 1ull);
        wk.result_index = (int)(idx + a.result_base);
        if (idx < a.cap_results) {
            ResultRec<2>& r = a.results[idx];
            const bool fin = st.final_step();
            bool touch = false;
            for (int d = 0; d < 2; ++d) {
                const dd A = fin ? st.pfA[d] : st.A[d], B = fin ? st.pfB[d] : st.B[d];
                dd nA; nA.hi = -A.hi; nA.lo = -A.lo;
                r.box[d][0] = dd_to_double(dd_add(B, nA));
                r.box[d][1] = dd_to_double(dd_add(B, A));
                dd nA2; nA2.hi = -st.A[d].hi; nA2.lo = -st.A[d].lo;
                const double lo = dd_to_double(dd_add(st.B[d], nA2));
                const double hi = dd_to_double(dd_add(st.B[d], st.A[d]));
                r.root[d] = fin ? (lo + hi) / 2 : (r.box[d][0] + r.box[d][1]) / 2;
                r.comb_iv[d][0] = fin ? st.pf_iv[d][0] : st.iv[d][0];
                r.comb_iv[d][1] = fin ? st.pf_iv[d][1] : st.iv[d][1];
                r.iv_lo[d] = st.iv[d][0];
                r.cell_iv[d][0] = wk.cell_iv[d][0]; r.cell_iv[d][1] = wk.cell_iv[d][1];
                r.next_mid[d] = st.next_mid[d];
                touch = touch || r.comb_iv[d][0] == wk.cell_iv[d][0] || r.comb_iv[d][1] == wk.cell_iv[d][1];
            }
            r.system = rec.system; r.parent_node = rec.parent_node; r.collector = rec.collector;
            r.level = wk.node_level;
            r.flags = extra_flags | (fin ? kResFinalStep : 0u) | (touch ? kResTouchesCell : 0u);
            r.pad = 0;
            r.path_hi = rec.path_hi; r.path_lo = rec.path_lo;
        } else at.add(&a.ctl->overflow, 1ull);
    };
    auto sizes = [&](int& T, int& n0max, int& n1max, unsigned long long& total) { box_sizes(rec, wk, T, n0max, n1max, total); };
    auto queue_restrict = [&]() {
        int T, n0m, n1m; unsigned long long total;
        sizes(T, n0m, n1m, total);
        wk.op = kTRestrict;
        dc.op = kTRestrict; pick_team(need_restrict_w(T, n0m, n1m), need_restrict(T, n0m, n1m), a.coarse, a.pol, 0, dc.need, dc.cls); dc.bound = total;
        dc.ext = (unsigned long long)(n0m > n1m ? n0m : n1m);
    };

    for (;;) {
        if (wk.fresh) {
            // ---- entry of one solvePolyRecursive invocation :1202-1241
            dc.boxes += 1;
            if (st.is_point()) { emit(kResPointAtEntry); break; }
            wk.node_level += 1;
            if ((unsigned long long)wk.node_level > dc.maxlevel) dc.maxlevel = wk.node_level;
            dc.coeffs += (unsigned long long)rec.shape[0][0] * rec.shape[0][1] + (unsigned long long)rec.shape[1][0] * rec.shape[1][1];
            bool discard = false;
            if (prm.constant_check)
                for (int p = 0; p < 2 && !discard; ++p) {
                    const double c0 = wk.low[p][0];
                    if (fabs(c0) > wk.sumabs[p] - fabs(c0) + rec.err[p]) { discard = true; dc.dconst += 1; }
                }
            if (!discard && (prm.low_dim_quad_check || prm.all_dim_quad_check))
                for (int p = 0; p < 2 && !discard; ++p)
                    if (quad_check_2d(wk.low[p], wk.sumabs[p], rec.err[p])) { discard = true; dc.dquad += 1; }
            if (discard) break;
            if (prm.trim) {                                          // trimMs :1232, computed by the producing tensor op
                bool any = false;
                for (int p = 0; p < 2; ++p) {
                    any = any || wk.tshape[p][0] != rec.shape[p][0] || wk.tshape[p][1] != rec.shape[p][1];
                    rec.shape[p][0] = wk.tshape[p][0]; rec.shape[p][1] = wk.tshape[p][1];
                    wk.sumabs[p] = wk.tsumabs[p]; rec.err[p] = wk.terr[p];
                }
                if (any) wk.trim_valid = 0;                         // a second trimMs of these tensors would start from here
            }
            for (int d = 0; d < 2; ++d) {
                wk.entry_iv[d][0] = st.final_step() ? st.pf_iv[d][0] : st.iv[d][0];
                wk.entry_iv[d][1] = st.final_step() ? st.pf_iv[d][1] : st.iv[d][1];
                wk.last_size[d] = st.iv[d][1] - st.iv[d][0];
            }
            wk.zoom_count = 0; wk.changed = 1; wk.should_stop = 0; wk.fresh = 0;
        } else {
            // after a restriction :952-954, :1249-1252
            if (st.final_step() && st.is_point()) { wk.should_stop = 1; wk.changed = 0; }
            bool all_big = true;
            for (int d = 0; d < 2; ++d) {
                const double now = st.iv[d][1] - st.iv[d][0];
                all_big = all_big && (now >= wk.last_size[d] / 2);
                wk.last_size[d] = now;
            }
            if (all_big) wk.zoom_count += 1;
        }
        // ---- zoom loop :1243-1252
        bool queued = false, dead = false;
        while (wk.changed && wk.zoom_count <= prm.max_zoom_count) {
            dc.zooms += 1;
            double lin[2][2], c0[2];
            for (int p = 0; p < 2; ++p) {
                c0[p] = wk.low[p][0];
                lin[p][0] = (rec.shape[p][0] == 1) ? 0.0 : wk.low[p][1];
                lin[p][1] = (rec.shape[p][1] == 1) ? 0.0 : wk.low[p][2];
            }
            double sub[2][2];
            BilsOut bo = bounding_interval<2>(lin, c0, wk.sumabs, rec.err, st.final_step(), sub);
            for (int d = 0; d < 2; ++d)                                       // frozen axes :932-935
                if (st.iv[d][0] == st.iv[d][1]) { sub[d][0] = -1.0; sub[d][1] = 1.0; }
            bool changed = bo.changed, should_stop = bo.should_stop, thr = bo.throw_out;
            if (thr && !st.can_throw_out()) { thr = false; should_stop = true; changed = true; }   // :937-940
            wk.changed = changed ? 1 : 0; wk.should_stop = should_stop ? 1 : 0;
            if (thr) { dc.dzoom += 1; dead = true; break; }
            if (changed) {
                double al[2], be[2];
                if (!add_transform<2>(st, sub, al, be)) { dc.dzoom += 1; dead = true; break; }
                for (int d = 0; d < 2; ++d) { wk.alpha[d] = al[d]; wk.beta[d] = be[d]; }
                queue_restrict();
                queued = true;
                break;
            }
            wk.zoom_count += 1;                                                // nothing changed: counts towards maxZoomCount (:1250)
        }
        if (queued || dead) break;
        // ---- what next? :1254-1317
        if (wk.should_stop) {
            if (st.final_step()) { emit(0u); break; }
            st.start_final_step();                                          // :1264 re-enter on the same tensors
            wk.fresh = 1;
            if (prm.trim && !wk.trim_valid) {
                // these tensors were trimmed at entry and not rewritten since: let a copy-only pass recompute
                // what a second trimMs does to them
                wk.alpha[0] = wk.alpha[1] = 1.0; wk.beta[0] = wk.beta[1] = 0.0; wk.copy_only = 1;
                queue_restrict();
                break;
            }
            continue;
        }
        if (st.final_step()) {
            st.flags |= kFlagCanThrowFinal;
            emit(kResCollector);                                            // the parent reports; its subtree supplies the points
            wk.final_split = 1;
        } else wk.final_split = 0;
        {
            int shape[2][2] = {{rec.shape[0][0], rec.shape[0][1]}, {rec.shape[1][0], rec.shape[1][1]}};
            int order[2][2];
            const int k = subdivision_axes<2>(shape, st.iv, wk.node_level, order);
            if (k == 0) { at.add(&a.ctl->no_axis, 1ull); break; }              // getInverseOrder of an empty order raises (:1010)
            wk.nsplit = k;
            for (int p = 0; p < 2; ++p) { wk.order[p][0] = order[p][0]; wk.order[p][1] = (k > 1) ? order[p][1] : order[p][0]; }
            wk.node_index = -1;
            if (!wk.final_split) {
                const unsigned long long ni = at.add(&a.ctl->n_nodes, 1ull);
                if (ni < a.cap_nodes) {
                    NodeRec<2>& nd = a.nodes[ni];
                    for (int d = 0; d < 2; ++d) { nd.iv[d][0] = wk.entry_iv[d][0]; nd.iv[d][1] = wk.entry_iv[d][1]; nd.next_mid[d] = st.next_mid[d]; }
                    nd.system = rec.system; nd.parent_node = rec.parent_node; nd.level = wk.node_level; nd.nsplit = k;
                    wk.node_index = (int)(ni + a.node_base);
                } else at.add(&a.ctl->overflow, 1ull);
            }
            int T, n0m, n1m; unsigned long long total;
            sizes(T, n0m, n1m, total);
            wk.op = kTSubdivide;
            dc.op = kTSubdivide; pick_team(need_subdivide_w(T), need_subdivide(T), a.coarse, a.pol, 1, dc.need, dc.cls); dc.bound = total << k;
            dc.ext = (unsigned long long)(n0m > n1m ? n0m : n1m);
            dc.subdiv += 1;
        }
        break;
    }

}

template <class At>
YR_HD void scalar_step(const Args& a, unsigned int b, Decision& dc, At at) { scalar_step_on(a, a.recs[b], a.work[b], dc, at); }

// Level-0 boxes arrive packed from yr_init_boxes: give them their Work state and queue a copy-only pass that
// re-lays the tensors out (odd row stride) and computes sums / low-order coefficients / trim outcome.
template <class At>
YR_HD void import_box(const Args& a, unsigned int b, Decision& dc, At at) {
    const BoxRec<2>& rec = a.recs[b];
    Work& w = a.work[b];
    w.off[0] = rec.data_off; w.off[1] = rec.data_off + (unsigned long long)rec.shape[0][0] * rec.shape[0][1];
    for (int p = 0; p < 2; ++p) {
        w.ld[p] = rec.shape[p][1];
        w.tshape[p][0] = rec.shape[p][0]; w.tshape[p][1] = rec.shape[p][1]; w.tsumabs[p] = 0.0; w.terr[p] = rec.err[p];
        w.sumabs[p] = 0.0;
        for (int q = 0; q < 6; ++q) w.low[p][q] = 0.0;
        w.alpha[p] = 1.0; w.beta[p] = 0.0; w.last_size[p] = 0.0;
        w.entry_iv[p][0] = w.entry_iv[p][1] = 0.0;
        w.cell_iv[p][0] = rec.st.iv[p][0]; w.cell_iv[p][1] = rec.st.iv[p][1];
        w.order[p][0] = w.order[p][1] = 0;
    }
    w.op = kTRestrict; w.fresh = 1; w.copy_only = 1; w.trim_valid = 0;
    w.zoom_count = 0; w.changed = 1; w.should_stop = 0; w.final_split = 0;
    w.node_level = rec.level; w.result_index = -1; w.node_index = -1; w.nsplit = 0;
    int T, n0m, n1m; unsigned long long total;
    box_sizes(rec, w, T, n0m, n1m, total);
    dc.op = kTRestrict; pick_team(need_restrict_w(T, n0m, n1m), need_restrict(T, n0m, n1m), a.coarse, a.pol, 0, dc.need, dc.cls); dc.bound = total;
    dc.ext = (unsigned long long)(n0m > n1m ? n0m : n1m);
    at.add(&a.ctl->n_boxes, 1ull);
}

// queue the box for its next tensor op (plain per-thread atomics; the CUDA backend aggregates per warp)
template <class At>
YR_HD void commit_decision(const Args& a, unsigned int b, const Decision& dc, At at) {
    if (dc.op == kTNone) return;
    const int slot = op_slot(dc.op);
    const unsigned long long pos = at.add(&a.ctl->n_list[slot][dc.cls], 1ull);
    a.lists[slot][dc.cls][pos] = b;
    at.max(&a.ctl->need_max[slot][dc.cls], dc.need);
    if (slot == 1) at.max(&a.ctl->ext_max[dc.cls], dc.ext);
    at.add(&a.ctl->out_bound, dc.bound);
}

}  // namespace f2
}  // namespace yr
