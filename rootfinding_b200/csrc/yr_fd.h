// Frontier pipeline for systems of D polynomials in D variables, D = 3, 4 ("fd"): the round structure, queues, size
// classes and control block of the 2-D pipeline (yr_f2.h) with tensor kernels written for D-dimensional tensors.
//
//   tensor ops   team (warp, or CTA of 4 / 8 warps) per box, one launch per (op, size class):
//                  RESTRICT   transformChebToInterval (:864-894): C(alpha_j, beta_j) per axis, every tensor restricted
//                             along every axis IN PLACE (fiber per lane, row blocks ascending -- see yr_f2.h)
//                  SUBDIVIDE  getSubdivisionIntervals (:1051-1129): the 2^k children of a box by recursive halving; at
//                             depth l the lower half goes to region l + 1, the upper half stays in place, so a box needs
//                             k + 1 tensor regions for any k
//                every produced tensor leaves with sum|M|, its coefficients of total degree <= 2 and the outcome of
//                trimMs (:1131-1171)
//   scalar step  thread per box (checks :1213-1224, BILS :628-753, addTransform :420-454, zoom bookkeeping :1243-1252,
//                final-step re-entry :1264, result / node records, next op): the D-dimensional mathematics of yr_math.h
//
// Tensor layout (arena and shared memory): row-major, last axis padded to an ODD stride ld, the axes above it tight:
// stride[D-1] = 1, stride[D-2] = ld, stride[j] = n[j+1] * stride[j+1].  A restriction shrinks an extent and keeps the
// strides (a view); tensors are re-laid out tightly when they leave for the arena.
//
// Arithmetic as in yr_f2.h: -fmad=false, matrix entries and accumulation order of TransformChebInPlace1D (:45-120), so
// restricted tensors are bit-identical to the reference's.
#pragma once
#include "yr_f2.h"

namespace yr {
namespace fd {

using f2::pair_t;
using f2::kOps; using f2::kTNone; using f2::kTRestrict; using f2::kTSubdivide; using f2::op_slot;
using f2::kClasses; using f2::kWarpClasses; using f2::kCta128Classes; using f2::class_threads;
using f2::kMaxNeed; using f2::kWarp32Max; using f2::TeamPolicy; using f2::Ctl; using f2::Tables; using f2::Decision;
using f2::odd_ld; using f2::even_up; using f2::panel_off; using f2::panel_doubles; using f2::PanelJob;

template <int D> constexpr int kLow = (D + 1) * (D + 2) / 2;      // coefficients of total degree <= 2 (== low_count(D))

template <int D>
struct Work {
    unsigned long long off[D];   // first double of tensor p in the arena the box currently lives in
    int32_t ld[D];               // its last-axis stride
    int32_t full[D][D];          // extents the strides of tensor p are built from (>= the current shape)
    int32_t tshape[D][D];        // extents trimMs would leave               } valid for the tensors as the last
    double  tsumabs[D], terr[D]; // sum|M_p| and error after that trim       } tensor op wrote them (trim_valid)
    double  sumabs[D];
    double  low[D][kLow<D>];   // [c0 | linear (D) | pure T2 (D) | cross pairs a < b]
    double  alpha[D], beta[D];
    double  entry_iv[D][2], cell_iv[D][2], last_size[D];
    int32_t op, fresh, copy_only, trim_valid;
    int32_t zoom_count, changed, should_stop, final_split;
    int32_t node_level, result_index, node_index, nsplit;
    int32_t order[D][D];
};

template <int D>
struct Args {
    BoxRec<D>* recs; Work<D>* work;
    const double* data_in; double* data_out; unsigned long long cap_data_out;
    ResultRec<D>* results; unsigned long long cap_results, result_base;
    NodeRec<D>* nodes; unsigned long long cap_nodes, node_base;
    unsigned int* lists[kOps][kClasses];
    unsigned long long cap_boxes;
    int coarse;
    TeamPolicy pol;
    Ctl* ctl;
    Tables tab;
    SolverParams prm;
};

// ---- tensor geometry -------------------------------------------------------------------------------
template <int D>
YR_HD void strides_of(const int* full, int ld, int* stride) {
    stride[D - 1] = 1;
    if (D >= 2) stride[D - 2] = ld;
    for (int j = D - 3; j >= 0; --j) stride[j] = full[j + 1] * stride[j + 1];
}
template <int D>
YR_HD int tensor_doubles(const int* full, int ld) {                // storage of a tensor with extents `full`
    int s = ld;
    for (int j = 0; j < D - 1; ++j) s *= full[j];
    return even_up(s);
}
template <int D>
YR_HD int tight_doubles(const int* shape) { return tensor_doubles<D>(shape, odd_ld(shape[D - 1])); }

// e / n for 0 <= e < 2^18, n >= 1: the element and fiber counts of one box.  On the device through a float reciprocal
// (4 instructions instead of the ~20 of an integer division): (e + 0.5) / n is at least 0.5 / n away from an integer,
// i.e. 0.5 / e >= 1.9e-6 relative, against the 2.4e-7 (2 ulp) of the fast division -- exact.
YR_HD int small_div(int e, int n) {
#ifdef __CUDA_ARCH__
    return __float2int_rz(__fdividef((float)e + 0.5f, (float)n));
#else
    return e / n;
#endif
}
// element e (row-major over `shape`) -> offset with `stride` (no division for the outermost axis)
template <int D>
YR_HD int view_offset(int e, const int* shape, const int* stride) {
    int off = 0;
#pragma unroll
    for (int a = D - 1; a >= 1; --a) { const int q = small_div(e, shape[a]); off += (e - q * shape[a]) * stride[a]; e = q; }
    return off + e * stride[0];
}
// fiber f along AXIS (row-major over the other axes) -> offset of its first element
template <int D, int AXIS>
YR_HD int fiber_offset_t(int f, const int* shape, const int* stride) {
    constexpr int kOuter = (AXIS == 0) ? 1 : 0;                 // the outermost axis that is not AXIS
    int off = 0;
#pragma unroll
    for (int a = D - 1; a > kOuter; --a) {
        if (a == AXIS) continue;
        const int q = small_div(f, shape[a]);
        off += (f - q * shape[a]) * stride[a];
        f = q;
    }
    return off + f * stride[kOuter];
}
template <int D>
YR_HD int fiber_offset(int f, int axis, const int* shape, const int* stride) {
    int off = 0;
    for (int a = D - 1; a >= 0; --a) {
        if (a == axis) continue;
        const int q = small_div(f, shape[a]);
        off += (f - q * shape[a]) * stride[a];
        f = q;
    }
    return off;
}
template <int D>
YR_HD int fibers_of(int axis, const int* shape) { int F = 1; for (int a = 0; a < D; ++a) if (a != axis) F *= shape[a]; return F; }

// ---- one restriction pass along AXIS, in place or into another region with the same strides --------------------
// (AXIS is a template parameter: with every index into shape / stride a compile-time constant the arrays live in registers)
template <int D, int AXIS, bool FMA, class Team>
YR_DEV double pass_axis_t(Team& tm, const double* src, double* dst, const int* shape, const int* stride, int r,
                          const double* Cp, int nt, unsigned long long& macs) {
    const int n = shape[AXIS], sk = stride[AXIS];
    int F = 1;
#pragma unroll
    for (int a = 0; a < D; ++a) if (a != AXIS) F *= shape[a];
    const int B = (r + 3) >> 2;
    double asum = 0.0;
    for (int f = tm.tid; f < F; f += tm.nthreads) {
        const int base = fiber_offset_t<D, AXIS>(f, shape, stride);
        const double* sc = src + base;
        double* dc = dst + base;
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
template <int D, bool FMA, class Team>
YR_DEV double pass_axis(Team& tm, const double* src, double* dst, const int* shape, const int* stride, int axis, int r,
                        const double* Cp, int nt, unsigned long long& macs) {
    switch (axis) {
        case 0: return pass_axis_t<D, 0, FMA>(tm, src, dst, shape, stride, r, Cp, nt, macs);
        case 1: return pass_axis_t<D, 1, FMA>(tm, src, dst, shape, stride, r, Cp, nt, macs);
        case 2: return pass_axis_t<D, 2, FMA>(tm, src, dst, shape, stride, r, Cp, nt, macs);
        default: return pass_axis_t<D, D - 1, FMA>(tm, src, dst, shape, stride, r, Cp, nt, macs);
    }
}

// both halves of a subdivision at m = 0 in one sweep (mirror-image matrices, see yr_f2.h): lower -> dstL, upper -> dstU;
// either may be src
template <int D, int AXIS, bool FMA, class Team>
YR_DEV void pass_axis_dual_t(Team& tm, const double* src, double* dstL, double* dstU, const int* shape, const int* stride,
                             int r, const double* Cp, int nt, unsigned long long& macs, double& sumL, double& sumU) {
    const int n = shape[AXIS], sk = stride[AXIS];
    int F = 1;
#pragma unroll
    for (int a = 0; a < D; ++a) if (a != AXIS) F *= shape[a];
    const int B = (r + 3) >> 2;
    double asl = 0.0, asu = 0.0;
    for (int f = tm.tid; f < F; f += tm.nthreads) {
        const int base = fiber_offset_t<D, AXIS>(f, shape, stride);
        for (int b = 0; b < B; ++b) {
            const int i0 = b << 2;
            YR_F2_DUAL_BLOCK(src + base + i0 * sk, dstL + base + i0 * sk, dstU + base + i0 * sk, sk, sk)
        }
    }
    if (tm.tid == 0) macs += 2ull * (unsigned)(F * (r * n - r * (r - 1) / 2));
    tm.sum2_sync(asl, asu, sumL, sumU);
}
template <int D, bool FMA, class Team>
YR_DEV void pass_axis_dual(Team& tm, const double* src, double* dstL, double* dstU, const int* shape, const int* stride, int axis,
                           int r, const double* Cp, int nt, unsigned long long& macs, double& sumL, double& sumU) {
    switch (axis) {
        case 0: pass_axis_dual_t<D, 0, FMA>(tm, src, dstL, dstU, shape, stride, r, Cp, nt, macs, sumL, sumU); break;
        case 1: pass_axis_dual_t<D, 1, FMA>(tm, src, dstL, dstU, shape, stride, r, Cp, nt, macs, sumL, sumU); break;
        case 2: pass_axis_dual_t<D, 2, FMA>(tm, src, dstL, dstU, shape, stride, r, Cp, nt, macs, sumL, sumU); break;
        default: pass_axis_dual_t<D, D - 1, FMA>(tm, src, dstL, dstU, shape, stride, r, Cp, nt, macs, sumL, sumU); break;
    }
}

template <int D, class Team>
YR_DEV double view_abs_sum(Team& tm, const double* M, const int* shape, const int* stride) {
    int cnt = 1;
    for (int a = 0; a < D; ++a) cnt *= shape[a];
    double s = 0.0;
    for (int e = tm.tid; e < cnt; e += tm.nthreads) s += fabs(M[view_offset<D>(e, shape, stride)]);
    return tm.sum(s);
}

// ---- what every produced tensor leaves behind --------------------------------------------------------------
template <int D>
struct Epilogue { double low[kLow<D>]; int t[D]; double tsum, terr; };

template <int D>
YR_HD double coef_or_zero(const double* M, const int* shape, const int* stride, int a, int va, int b, int vb) {
    int off = 0;
    if (a >= 0) { if (va >= shape[a]) return 0.0; off += va * stride[a]; }
    if (b >= 0) { if (vb >= shape[b]) return 0.0; off += vb * stride[b]; }
    return M[off];
}

// abs-sum of the hyperplane idx[AXIS] = j of a view with extents t, for a team of one warp: lanes take ROWS of the plane
// (runs along the innermost axis that is not AXIS), one index decode per row instead of one per element.  (Measured: the
// 4-D subdivide kernel spent 45 % of its instructions in the per-element decode; for the 128-thread CTA teams of the 3-D
// boxes the per-element loop below stays faster -- too few rows for 128 threads.  profiles/r02_fd.txt)
template <int D, int AXIS, class Team>
YR_DEV double plane_abs_sum_rows(Team& tm, const double* M, const int* t, const int* stride, int j) {
    constexpr int INNER = (AXIS == D - 1) ? D - 2 : D - 1;
    constexpr int kOuter = (AXIS == 0 || INNER == 0) ? ((AXIS == 1 || INNER == 1) ? 2 : 1) : 0;   // outermost remaining axis
    int rows = 1;
#pragma unroll
    for (int a = 0; a < D; ++a) if (a != AXIS && a != INNER) rows *= t[a];
    const int len = t[INNER], si = stride[INNER];
    const double* plane = M + j * stride[AXIS];
    double s = 0.0;
    for (int r = tm.tid; r < rows; r += tm.nthreads) {
        int rem = r, off = 0;
#pragma unroll
        for (int a = D - 1; a >= 0; --a) {
            if (a == AXIS || a == INNER) continue;
            if (a == kOuter) off += rem * stride[a];
            else { const int q = small_div(rem, t[a]); off += (rem - q * t[a]) * stride[a]; rem = q; }
        }
        const double* row = plane + off;
        for (int i = 0; i < len; ++i) s += fabs(row[i * si]);
    }
    return tm.sum(s);
}

// Collective over the whole team; every thread gets the results.
template <int D, class Team>
YR_DEV void tensor_epilogue(Team& tm, const double* M, const int* shape, const int* stride, double sumabs, double err,
                            const SolverParams& prm, Epilogue<D>& ep) {
    int q = 0;
    ep.low[q++] = M[0];
    for (int a = 0; a < D; ++a) ep.low[q++] = coef_or_zero<D>(M, shape, stride, a, 1, -1, 0);
    for (int a = 0; a < D; ++a) ep.low[q++] = coef_or_zero<D>(M, shape, stride, a, 2, -1, 0);
    for (int a = 0; a < D; ++a) for (int b = a + 1; b < D; ++b) ep.low[q++] = coef_or_zero<D>(M, shape, stride, a, 1, b, 1);
    int t[D];
    for (int a = 0; a < D; ++a) t[a] = shape[a];
    double budget = prm.trim_abs_tol + err * prm.trim_rel_tol;
    double removed = 0.0;
    if (prm.trim) {
        for (int axis = 0; axis < D; ++axis) {
            for (;;) {
                if (t[axis] <= 3) break;
                // abs-sum of the last hyperplane over the CURRENT (already trimmed) extents of the other axes
                int cnt = 1;
                for (int a = 0; a < D; ++a) if (a != axis) cnt *= t[a];
                const double* plane = M + (t[axis] - 1) * stride[axis];
                double s = 0.0;
                if (tm.nthreads <= 32 && D >= 4) {
                    switch (axis) {
                        case 0: s = plane_abs_sum_rows<D, 0>(tm, M, t, stride, t[axis] - 1); break;
                        case 1: s = plane_abs_sum_rows<D, 1>(tm, M, t, stride, t[axis] - 1); break;
                        case 2: s = plane_abs_sum_rows<D, 2>(tm, M, t, stride, t[axis] - 1); break;
                        default: s = plane_abs_sum_rows<D, D - 1>(tm, M, t, stride, t[axis] - 1); break;
                    }
                } else {
                    for (int e = tm.tid; e < cnt; e += tm.nthreads) s += fabs(plane[fiber_offset<D>(e, axis, t, stride)]);
                    s = tm.sum(s);
                }
                if (s < budget) { budget -= s; err += s; removed += s; t[axis] -= 1; } else break;
            }
        }
    }
    for (int a = 0; a < D; ++a) ep.t[a] = t[a];
    ep.terr = err;
    ep.tsum = sumabs - removed;
}

template <int D>
YR_HD void store_epilogue(Work<D>& w, int p, const Epilogue<D>& ep, double sumabs) {
    w.sumabs[p] = sumabs;
    for (int q = 0; q < kLow<D>; ++q) w.low[p][q] = ep.low[q];
    for (int a = 0; a < D; ++a) w.tshape[p][a] = ep.t[a];
    w.tsumabs[p] = ep.tsum; w.terr[p] = ep.terr;
}

// view (shape, stride) in shared memory -> arena, tight with last-axis stride ld_out.  When the view still has the
// extents its strides were built from (`full`) and the same last-axis stride, the arena image is the shared-memory image:
// one flat copy (the common case for the children of a subdivision).  Otherwise row by row: one decode per row.
template <int D, class Team>
YR_DEV void copy_out(Team& tm, double* dst, int ld_out, const double* M, const int* shape, const int* full, const int* stride) {
    bool same = (ld_out == stride[D - 2 >= 0 ? D - 2 : 0]);
#pragma unroll
    for (int a = 1; a < D; ++a) same = same && (shape[a] == full[a]);
    if (same) {
        const int n2 = tensor_doubles<D>(shape, ld_out) >> 1;
        const pair_t* s2 = reinterpret_cast<const pair_t*>(M);
        pair_t* d2 = reinterpret_cast<pair_t*>(dst);
        for (int e = tm.tid; e < n2; e += tm.nthreads) d2[e] = s2[e];
        return;
    }
    int ostr[D], rshape[D];
    strides_of<D>(shape, ld_out, ostr);
    int rows = 1;
#pragma unroll
    for (int a = 0; a < D - 1; ++a) { rows *= shape[a]; rshape[a] = shape[a]; }
    rshape[D - 1] = 1;
    const int len = shape[D - 1];
    // four threads per row (rows are short), teams are powers of two
    const int cg = tm.nthreads >= 4 ? 4 : tm.nthreads;
    const int per = tm.nthreads / cg, r0 = tm.tid / cg, j0 = tm.tid - r0 * cg;
    for (int r = r0; r < rows; r += per) {
        const double* srow = M + view_offset<D>(r, rshape, stride);
        double* drow = dst + view_offset<D>(r, rshape, ostr);
        for (int j = j0; j < len; j += cg) drow[j] = srow[j];
    }
}

// arena -> shared memory: flat when the arena layout can be used as it is (odd ld, even offset), else re-laid out
// (level-0 boxes arrive packed with ld = n[D-1])
template <int D, class Team>
YR_DEV void fetch(Team& tm, double* dst, const int* full, int lds, const double* src, int ldg, bool flat) {
    if (flat) { tm.copy_async(dst, src, tensor_doubles<D>(full, lds)); return; }
    int sstr[D], dstr[D];
    strides_of<D>(full, ldg, sstr); strides_of<D>(full, lds, dstr);
    int cnt = 1;
    for (int a = 0; a < D; ++a) cnt *= full[a];
    for (int e = tm.tid; e < cnt; e += tm.nthreads) dst[view_offset<D>(e, full, dstr)] = src[view_offset<D>(e, full, sstr)];
}

// ---- shared-memory needs ---------------------------------------------------------------------------
template <int D>
YR_HD unsigned long long need_restrict(int T, const int* nmax) {     // [A: T][C_0 .. C_{D-1}][rows_after ints]
    unsigned long long need = (unsigned long long)T + 2;
    int rows = 0;
    for (int a = 0; a < D; ++a) { need += (unsigned)panel_doubles(nmax[a]); rows += nmax[a]; }
    return need + (unsigned)((rows + 1) / 2);
}
YR_HD unsigned long long need_subdivide(int T, int k) { return (unsigned long long)(k + 1) * T; }
YR_HD void pick_team(unsigned long long need, int coarse, const TeamPolicy& pol, int slot, int& cls) {
    if (need <= pol.warp_max[slot]) {
        cls = f2::warp_class(need);
        if (cls < f2::kLane16Classes) cls = f2::kLane16Classes;        // no 16-lane teams here: the smallest warp class
        if (coarse) cls = kWarpClasses - 1;
    } else {
        cls = f2::cta_class(need, pol.cta128_max);
        if (coarse) cls = cls < kCta128Classes ? kCta128Classes - 1 : kClasses - 1;
    }
}

template <int D>
YR_HD void box_sizes(const BoxRec<D>& rec, const Work<D>& wk, int& T, int* nmax, unsigned long long& total) {
    T = 0; total = 0;
    for (int a = 0; a < D; ++a) nmax[a] = 0;
    for (int p = 0; p < D; ++p) {
        const int l = odd_ld(wk.full[p][D - 1]) > wk.ld[p] ? odd_ld(wk.full[p][D - 1]) : wk.ld[p];
        const int s = tensor_doubles<D>(wk.full[p], l);
        if (s > T) T = s;
        total += (unsigned long long)s;
        for (int a = 0; a < D; ++a) if (rec.shape[p][a] > nmax[a]) nmax[a] = rec.shape[p][a];
    }
}

// ---- RESTRICT (any team) --------------------------------------------------------------------------------
// slice: [A: T][C_0 .. C_{D-1}][rows_after ints]
template <int D, bool FMA, class Team>
YR_DEV void restrict_box(Team& tm, const Args<D>& a, unsigned int b, double* slice, unsigned long long& macs) {
    BoxRec<D>& rec = a.recs[b];
    Work<D>& wk = a.work[b];
    const SolverParams& prm = a.prm;
    const int copy_only = wk.copy_only;
    // geometry of tensor p, straight from the records into registers.  (Kept in arrays indexed by the run-time p of the
    // un-unrolled tensor loop below they lived in local memory: long-scoreboard waits, profiles/r02_fd.txt.  Thread 0
    // rewrites the entries of tensor p only after every thread has read them.)
    auto geom = [&](int p, int* shp_p, int* full_p, int& ldg_p, int& lds_p, unsigned long long& off_p, bool& flat_p) {
#pragma unroll
        for (int ax = 0; ax < D; ++ax) { shp_p[ax] = rec.shape[p][ax]; full_p[ax] = wk.full[p][ax]; }
        ldg_p = wk.ld[p]; off_p = wk.off[p];
        flat_p = (ldg_p & 1) && ldg_p >= full_p[D - 1] && (off_p & 1ull) == 0;
        lds_p = flat_p ? ldg_p : odd_ld(full_p[D - 1]);
    };
    int nmax[D];
    int T = 0;
    for (int ax = 0; ax < D; ++ax) nmax[ax] = 0;
    unsigned long long out_total = 0;
#pragma unroll
    for (int p = 0; p < D; ++p) {
        int shp_p[D], full_p[D], ldg_p, lds_p; unsigned long long off_p; bool flat_p;
        geom(p, shp_p, full_p, ldg_p, lds_p, off_p, flat_p);
#pragma unroll
        for (int ax = 0; ax < D; ++ax) if (shp_p[ax] > nmax[ax]) nmax[ax] = shp_p[ax];
        const int sz_p = tensor_doubles<D>(full_p, lds_p);
        if (sz_p > T) T = sz_p;
        out_total += (unsigned long long)tight_doubles<D>(shp_p);
    }
    double* A = slice;
    double* C[D]; int* rows[D];
    {
        double* c = slice + T;
        for (int ax = 0; ax < D; ++ax) { C[ax] = c; c += panel_doubles(nmax[ax]); }
        int* r = reinterpret_cast<int*>(c);
        for (int ax = 0; ax < D; ++ax) { rows[ax] = r; r += nmax[ax]; }
    }
    double al[D], be[D]; bool ident[D];
    for (int ax = 0; ax < D; ++ax) { al[ax] = wk.alpha[ax]; be[ax] = wk.beta[ax]; ident[ax] = copy_only || (al[ax] == 1.0 && be[ax] == 0.0); }
    unsigned long long base = 0;
    if (tm.tid == 0) base = tm.atomic_add(&a.ctl->data_used, out_total);
    base = tm.bcast(base);
    if (base + out_total > a.cap_data_out) { if (tm.tid == 0) tm.atomic_add(&a.ctl->overflow, 1ull); return; }
    tm.sync();                                                   // the previous box is done with the slice
    {
        int shp_p[D], full_p[D], ldg_p, lds_p; unsigned long long off_p; bool flat_p;
        geom(0, shp_p, full_p, ldg_p, lds_p, off_p, flat_p);
        fetch<D>(tm, A, full_p, lds_p, a.data_in + off_p, ldg_p, flat_p);               // runs behind the matrix build
    }
    for (int ax = 0; ax < D; ax += 2) {
        const int ax1 = ax + 1 < D ? ax + 1 : ax;
        PanelJob j0 = {C[ax], rows[ax], (!ident[ax] && nmax[ax] > 1) ? nmax[ax] : 0, al[ax], be[ax]};
        PanelJob j1 = {C[ax1], rows[ax1], (ax1 != ax && !ident[ax1] && nmax[ax1] > 1) ? nmax[ax1] : 0, al[ax1], be[ax1]};
        if (j0.nt > 0 || j1.nt > 0) f2::build_two(tm, j0, j1);
    }
    unsigned long long out_off = base;
#pragma unroll 1
    for (int p = 0; p < D; ++p) {
        int stride[D], cur[D], full_p[D], ldg_p, lds_p; unsigned long long off_p; bool flat_p;
        geom(p, cur, full_p, ldg_p, lds_p, off_p, flat_p);
        const double sum_in = wk.sumabs[p], err_in = rec.err[p];
        strides_of<D>(full_p, lds_p, stride);
        if (p > 0) fetch<D>(tm, A, full_p, lds_p, a.data_in + off_p, ldg_p, flat_p);
        tm.copy_wait();
        tm.sync();
        double s = copy_only ? view_abs_sum<D>(tm, A, cur, stride) : sum_in;
        double err = err_in;
#pragma unroll
        for (int axis = 0; axis < D; ++axis) {                       // unrolled: the axis of every pass is a compile-time constant
            const int n = cur[axis];
            if (!copy_only) err += n * kMachEps * s;                 // charged before the axis (:860)
            if (ident[axis] || n <= 1) continue;
            const int r = rows[axis][n - 1];
            s = pass_axis<D, FMA>(tm, A, A, cur, stride, axis, r, C[axis], nmax[axis], macs);
            cur[axis] = r;
        }
        const int ld_out = odd_ld(cur[D - 1]);
        const int osz = tight_doubles<D>(cur);
        {
            Epilogue<D> ep;
            tensor_epilogue<D>(tm, A, cur, stride, s, err, prm, ep);
            if (tm.tid == 0) {
                store_epilogue<D>(wk, p, ep, s);
                for (int ax = 0; ax < D; ++ax) { rec.shape[p][ax] = cur[ax]; wk.full[p][ax] = cur[ax]; }
                rec.err[p] = err;
                wk.off[p] = out_off; wk.ld[p] = ld_out;
            }
        }
        copy_out<D>(tm, a.data_out + out_off, ld_out, A, cur, full_p, stride);
        out_off += (unsigned long long)osz;
        tm.sync();                                               // tensor p has left A before tensor p + 1 arrives
    }
    if (tm.tid == 0) { wk.op = kTNone; wk.copy_only = 0; wk.trim_valid = 1; }
}

// ---- SUBDIVIDE (any team) ---------------------------------------------------------------------------------
template <int D>
YR_HD void child_record(const Args<D>& a, const BoxRec<D>& rec, const Work<D>& wk, unsigned int child, int c, int k, const int* split_axes) {
    // :1111-1128 -- thread-serial
    BoxRec<D>& out = a.recs[child];
    Work<D>& ow = a.work[child];
    IntervalState<D> st = rec.st;
    for (int j = 0; j < k; ++j) {
        const int ax = split_axes[j];
        const int half = (c >> (k - 1 - j)) & 1;
        const double mid = st.next_mid[ax];
        double sub[D][2], al[D], be[D];
        for (int d = 0; d < D; ++d) { sub[d][0] = -1.0; sub[d][1] = 1.0; }
        if (half) sub[ax][0] = mid; else sub[ax][1] = mid;
        add_transform<D>(st, sub, al, be);
    }
    for (int j = 0; j < k; ++j) st.next_mid[split_axes[j]] = 0.0;
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
    for (int d = 0; d < D; ++d) { ow.cell_iv[d][0] = st.iv[d][0]; ow.cell_iv[d][1] = st.iv[d][1]; }
    ow.op = kTNone; ow.fresh = 1; ow.copy_only = 0; ow.trim_valid = 1;
    ow.zoom_count = 0; ow.changed = 1; ow.should_stop = 0; ow.final_split = 0;
    ow.node_level = wk.node_level; ow.result_index = -1; ow.node_index = -1; ow.nsplit = 0;
}

// slice: k + 1 regions of T doubles; matrices from the level-independent tables (shared-memory copy for m = 0).
// The 2^k children are produced by a depth-first walk over the halvings (order[p][0], order[p][1], ...): at depth l the
// lower half is written to region l + 1 and the upper half replaces the tensor where it stands, so the walk first
// finishes everything below the lower half and then continues from the upper half with region l + 1 free again.
// State of the depth-first walk, ONE copy per warp in memory every lane of the warp can read (shared memory on the
// device; the warps of a 128-thread CTA team each keep their own identical copy, so a step costs a warp barrier, not a
// CTA barrier; 256-thread CTAs share one copy: their static shared memory is tight against the largest box).  As per-thread arrays indexed by the run-time depth they live in local memory: 400 bytes x every resident
// thread does not fit L1 next to the shared-memory carve-out, and a third of the subdivide kernels' stall samples were
// long-scoreboard waits on them (profiles/r02_fd.txt).  Every lane executes the walk and writes the SAME values; one
// warp barrier per step separates a step's reads from the next step's writes.
template <int D>
struct WalkState {
    double sm[D + 1], er[D + 1], upper_sum[D + 1], child_err[D + 1];
    int cur[D + 1], phase[D + 1];
    int shp[D + 1][D], upper_shape[D + 1][D];
};

template <int D, bool FMA, class Team>
YR_DEV void subdivide_box(Team& tm, const Args<D>& a, unsigned int b, double* slice, const double* smat_lo, const double* smat_hi,
                          int smat_nt, unsigned long long& macs, WalkState<D>* W, bool state_per_warp = true) {
    const BoxRec<D>& rec = a.recs[b];
    const Work<D>& wk = a.work[b];
    const SolverParams& prm = a.prm;
    const int k = wk.nsplit, nchild = 1 << k;
    int T = 0;
    unsigned long long per_child = 0;
#pragma unroll
    for (int p = 0; p < D; ++p) { const int s = tensor_doubles<D>(wk.full[p], wk.ld[p]); if (s > T) T = s; per_child += (unsigned long long)s; }
    int split_axes[D], midmask = 0;
    { bool used[D]; for (int d = 0; d < D; ++d) used[d] = false; for (int j = 0; j < k; ++j) used[wk.order[0][j]] = true; int j = 0; for (int ax = 0; ax < D; ++ax) if (used[ax]) split_axes[j++] = ax; for (; j < D; ++j) split_axes[j] = split_axes[0]; }
    for (int ax = 0; ax < D; ++ax) midmask |= ((rec.st.next_mid[ax] == 0.0) ? 0 : 1) << ax;
    int usedmask = 0;                                             // (registers: the walk indexes no per-thread array by depth)
    for (int j = 0; j < k; ++j) usedmask |= 1 << wk.order[0][j];
    unsigned long long cbase = 0, dbase = 0;
    if (tm.tid == 0) {
        cbase = tm.atomic_add(&a.ctl->n_boxes, (unsigned long long)nchild);
        dbase = tm.atomic_add(&a.ctl->data_used, per_child * nchild);
        for (int j = 0; j < k; ++j) { const double m = rec.st.next_mid[split_axes[j]]; if (m != 0.0 && m != kFirstSplit) tm.atomic_add(&a.ctl->bad_mid, 1ull); }
    }
    cbase = tm.bcast(cbase); dbase = tm.bcast(dbase);
    if (cbase + nchild > a.cap_boxes || dbase + per_child * nchild > a.cap_data_out) { if (tm.tid == 0) tm.atomic_add(&a.ctl->overflow, 1ull); return; }
    tm.sync();
    tm.copy_async(slice, a.data_in + wk.off[0], tensor_doubles<D>(wk.full[0], wk.ld[0]));   // tensor 0 arrives while the child records are written
    for (int c = tm.tid; c < nchild; c += tm.nthreads) child_record<D>(a, rec, wk, (unsigned int)(cbase + c), c, k, split_axes);
    unsigned dead = 0;                                            // children already known to fail the constant check
    unsigned long long poff = 0;
    for (int p = 0; p < D; ++p) {
        const int sz_p = tensor_doubles<D>(wk.full[p], wk.ld[p]);
        int ordp = 0;                                               // the halving order of this tensor, 4 bits per depth
        for (int j = 0; j < k; ++j) ordp |= wk.order[p][j] << (4 * j);
        if (p > 0) { tm.sync(); tm.copy_async(slice, a.data_in + wk.off[p], sz_p); }
        tm.copy_wait();
        tm.sync();
        int stride[D], fullp[D];
        for (int d = 0; d < D; ++d) fullp[d] = wk.full[p][d];        // (registers: `wk` is global memory the stores below may alias)
        strides_of<D>(fullp, wk.ld[p], stride);
        // depth-first walk: level l holds the tensor in region cur[l] with view shp[l], sum|M| sm[l], error er[l]
        int (&cur)[D + 1] = W->cur; int (&phase)[D + 1] = W->phase;
        int (&shp)[D + 1][D] = W->shp; int (&upper_shape)[D + 1][D] = W->upper_shape;
        double (&sm)[D + 1] = W->sm; double (&er)[D + 1] = W->er; double (&upper_sum)[D + 1] = W->upper_sum; double (&child_err)[D + 1] = W->child_err;
        for (int d = 0; d < D; ++d) shp[0][d] = rec.shape[p][d];
        cur[0] = 0; sm[0] = wk.sumabs[p]; er[0] = rec.err[p]; phase[0] = 0;
        int l = 0, halfmask = 0;
        while (l >= 0) {
            const int ph = l < k ? phase[l] : -1;
            // every lane sharing this copy of the state has read the step's phase, the previous step's writes are done
            if (state_per_warp) tm.wsync(); else tm.sync();
            if (l == k) {
                // ---- a finished child tensor leaves for the arena
                int c = 0;                                              // split axes ascending, smallest most significant
#pragma unroll
                for (int ax = 0; ax < D; ++ax) if ((usedmask >> ax) & 1) c = (c << 1) | ((halfmask >> ax) & 1);
                const double* M = slice + (size_t)cur[l] * T;
                const int ld_out = odd_ld(shp[l][D - 1]);
                const unsigned long long off = dbase + per_child * c + poff;
                Epilogue<D> ep;
                // a child the constant check (:1213) is going to discard needs neither its trim outcome nor its tensors (yr_f2.h)
                const double c0 = M[0];
                if (YR_F2_SKIP_DEAD && prm.constant_check && (fabs(c0) > sm[l] - fabs(c0) + er[l])) dead |= 1u << c;
                const bool is_dead = (dead >> c) & 1u;
                if (!is_dead) {
                    tensor_epilogue<D>(tm, M, shp[l], stride, sm[l], er[l], prm, ep);
                    copy_out<D>(tm, a.data_out + off, ld_out, M, shp[l], fullp, stride);
                } else {
                    for (int q = 0; q < kLow<D>; ++q) ep.low[q] = 0.0;
                    ep.low[0] = c0; ep.tsum = sm[l]; ep.terr = er[l];
                    for (int d = 0; d < D; ++d) ep.t[d] = shp[l][d];
                }
                if (tm.tid == 0) {
                    const unsigned int child = (unsigned int)(cbase + c);
                    Work<D>& ow = a.work[child]; BoxRec<D>& orec = a.recs[child];
                    store_epilogue<D>(ow, p, ep, sm[l]);
                    for (int d = 0; d < D; ++d) { orec.shape[p][d] = shp[l][d]; ow.full[p][d] = shp[l][d]; }
                    orec.err[p] = er[l];
                    ow.off[p] = off; ow.ld[p] = ld_out;
                }
                --l;
                continue;
            }
            const int axis = (ordp >> (4 * l)) & 15;
            if (ph == 0) {
                // ---- halve along `axis`: lower -> region l + 1, upper in place
                const int n = shp[l][axis];
                const int ms = (midmask >> axis) & 1;
                const bool smem = (ms == 0 && smat_lo != nullptr);
                const double* mlo = smem ? smat_lo : a.tab.mat[ms][0];
                const double* mhi = smem ? smat_hi : a.tab.mat[ms][1];
                const int mnt = smem ? smat_nt : a.tab.nt;
                const double e = n * kMachEps * sm[l];
                child_err[l] = e + er[l];
                double* X = slice + (size_t)cur[l] * T; double* L = slice + (size_t)(l + 1) * T;
                for (int d = 0; d < D; ++d) { shp[l + 1][d] = shp[l][d]; upper_shape[l][d] = shp[l][d]; }
                double sumL, sumU;
                tm.sync();                                           // region l + 1 has been left by the previous subtree
                if (n > 1) {
                    const int rl = a.tab.rows_after[ms][0][n - 1], ru = a.tab.rows_after[ms][1][n - 1];
                    shp[l + 1][axis] = rl; upper_shape[l][axis] = ru;
                    if (ms == 0 && rl == ru) pass_axis_dual<D, FMA>(tm, X, L, X, shp[l], stride, axis, rl, mlo, mnt, macs, sumL, sumU);
                    else {
                        sumL = pass_axis<D, FMA>(tm, X, L, shp[l], stride, axis, rl, mlo, mnt, macs);
                        sumU = pass_axis<D, FMA>(tm, X, X, shp[l], stride, axis, ru, mhi, mnt, macs);
                    }
                } else {
                    // extent-1 axis: TransformChebInPlaceND returns the tensor unchanged (:363)
                    int cnt = 1;
                    for (int d = 0; d < D; ++d) cnt *= shp[l][d];
                    for (int q = tm.tid; q < cnt; q += tm.nthreads) { const int o = view_offset<D>(q, shp[l], stride); L[o] = X[o]; }
                    sumL = sumU = sm[l];
                    tm.sync();
                }
                upper_sum[l] = sumU;
                phase[l] = 1;
                halfmask &= ~(1 << axis);
                cur[l + 1] = l + 1; sm[l + 1] = sumL; er[l + 1] = child_err[l]; phase[l + 1] = 0;
                ++l;
            } else if (ph == 1) {
                phase[l] = 2;
                halfmask |= 1 << axis;
                cur[l + 1] = cur[l]; sm[l + 1] = upper_sum[l]; er[l + 1] = child_err[l]; phase[l + 1] = 0;
                for (int d = 0; d < D; ++d) shp[l + 1][d] = upper_shape[l][d];
                ++l;
            } else --l;
        }
        poff += (unsigned long long)sz_p;
    }
    tm.sync();
}

// ---- scalar step (thread per box) -------------------------------------------------------------------
template <int D>
YR_HDM bool quad_check_low(const double* low, double sumabs, double tol) {
    // low = [c0 | linear (D) | pure (D) | cross pairs a < b]
    if (D == 3) {
        const double c[10] = {low[0], low[1], low[2], low[3], low[7], low[8], low[9], low[4], low[5], low[6]};
        return quad_check_3d(c, sumabs, tol);
    }
    double g[D], pure[D], cross[D][D];
    double used = fabs(low[0]);
    for (int a = 0; a < D; ++a) { g[a] = low[1 + a]; pure[a] = low[1 + D + a]; used += fabs(g[a]) + fabs(pure[a]); }
    int q = 1 + 2 * D;
    for (int a = 0; a < D; ++a) for (int b = 0; b < D; ++b) cross[a][b] = 0.0;
    for (int a = 0; a < D; ++a) for (int b = a + 1; b < D; ++b) { cross[a][b] = low[q++]; used += fabs(cross[a][b]); }
    return quad_check_nd<D>(low[0], g, pure, cross, sumabs - used, tol);
}

template <int D, class At>
YR_HDM void scalar_step(const Args<D>& a, unsigned int b, Decision& dc, At at) {
    BoxRec<D>& rec = a.recs[b];
    Work<D>& wk = a.work[b];
    const SolverParams& prm = a.prm;
    IntervalState<D>& st = rec.st;

    auto emit = [&](uint32_t extra_flags) {
        const unsigned long long idx = at.add(&a.ctl->n_results, 1ull);
        wk.result_index = (int)(idx + a.result_base);
        if (idx < a.cap_results) {
            ResultRec<D>& r = a.results[idx];
            const bool fin = st.final_step();
            bool touch = false;
            for (int d = 0; d < D; ++d) {
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
    auto queue_restrict = [&]() {
        int T, nmax[D]; unsigned long long total;
        box_sizes<D>(rec, wk, T, nmax, total);
        wk.op = kTRestrict;
        dc.op = kTRestrict; dc.need = need_restrict<D>(T, nmax); pick_team(dc.need, a.coarse, a.pol, 0, dc.cls); dc.bound = total;
        int e = 0; for (int d = 0; d < D; ++d) if (nmax[d] > e) e = nmax[d];
        dc.ext = (unsigned long long)e;
    };

    for (;;) {
        if (wk.fresh) {
            // ---- entry of one solvePolyRecursive invocation :1202-1241
            dc.boxes += 1;
            if (st.is_point()) { emit(kResPointAtEntry); break; }
            wk.node_level += 1;
            if ((unsigned long long)wk.node_level > dc.maxlevel) dc.maxlevel = wk.node_level;
            for (int p = 0; p < D; ++p) { unsigned long long c = 1; for (int d = 0; d < D; ++d) c *= (unsigned long long)rec.shape[p][d]; dc.coeffs += c; }
            bool discard = false;
            if (prm.constant_check)
                for (int p = 0; p < D && !discard; ++p) {
                    const double c0 = wk.low[p][0];
                    if (fabs(c0) > wk.sumabs[p] - fabs(c0) + rec.err[p]) { discard = true; dc.dconst += 1; }
                }
            if (!discard && ((prm.low_dim_quad_check && D <= 3) || prm.all_dim_quad_check))
                for (int p = 0; p < D && !discard; ++p)
                    if (quad_check_low<D>(wk.low[p], wk.sumabs[p], rec.err[p])) { discard = true; dc.dquad += 1; }
            if (discard) break;
            if (prm.trim) {                                          // trimMs :1232, computed by the producing tensor op
                bool any = false;
                for (int p = 0; p < D; ++p) {
                    for (int d = 0; d < D; ++d) { any = any || wk.tshape[p][d] != rec.shape[p][d]; rec.shape[p][d] = wk.tshape[p][d]; }
                    wk.sumabs[p] = wk.tsumabs[p]; rec.err[p] = wk.terr[p];
                }
                if (any) wk.trim_valid = 0;
            }
            for (int d = 0; d < D; ++d) {
                wk.entry_iv[d][0] = st.final_step() ? st.pf_iv[d][0] : st.iv[d][0];
                wk.entry_iv[d][1] = st.final_step() ? st.pf_iv[d][1] : st.iv[d][1];
                wk.last_size[d] = st.iv[d][1] - st.iv[d][0];
            }
            wk.zoom_count = 0; wk.changed = 1; wk.should_stop = 0; wk.fresh = 0;
        } else {
            // after a restriction :952-954, :1249-1252
            if (st.final_step() && st.is_point()) { wk.should_stop = 1; wk.changed = 0; }
            bool all_big = true;
            for (int d = 0; d < D; ++d) {
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
            double lin[D][D], c0[D];
            for (int p = 0; p < D; ++p) {
                c0[p] = wk.low[p][0];
                for (int d = 0; d < D; ++d) lin[p][d] = (rec.shape[p][d] == 1) ? 0.0 : wk.low[p][1 + d];
            }
            double sub[D][2];
            BilsOut bo = bounding_interval<D>(lin, c0, wk.sumabs, rec.err, st.final_step(), sub);
            for (int d = 0; d < D; ++d)                                       // frozen axes :932-935
                if (st.iv[d][0] == st.iv[d][1]) { sub[d][0] = -1.0; sub[d][1] = 1.0; }
            bool changed = bo.changed, should_stop = bo.should_stop, thr = bo.throw_out;
            if (thr && !st.can_throw_out()) { thr = false; should_stop = true; changed = true; }   // :937-940
            wk.changed = changed ? 1 : 0; wk.should_stop = should_stop ? 1 : 0;
            if (thr) { dc.dzoom += 1; dead = true; break; }
            if (changed) {
                double al[D], be[D];
                if (!add_transform<D>(st, sub, al, be)) { dc.dzoom += 1; dead = true; break; }
                for (int d = 0; d < D; ++d) { wk.alpha[d] = al[d]; wk.beta[d] = be[d]; }
                queue_restrict();
                queued = true;
                break;
            }
            wk.zoom_count += 1;
        }
        if (queued || dead) break;
        // ---- what next? :1254-1317
        if (wk.should_stop) {
            if (st.final_step()) { emit(0u); break; }
            st.start_final_step();                                          // :1264 re-enter on the same tensors
            wk.fresh = 1;
            if (prm.trim && !wk.trim_valid) {
                for (int d = 0; d < D; ++d) { wk.alpha[d] = 1.0; wk.beta[d] = 0.0; }
                wk.copy_only = 1;
                queue_restrict();
                break;
            }
            continue;
        }
        if (st.final_step()) {
            st.flags |= kFlagCanThrowFinal;
            emit(kResCollector);
            wk.final_split = 1;
        } else wk.final_split = 0;
        {
            int shape[D][D], order[D][D];
            for (int p = 0; p < D; ++p) for (int d = 0; d < D; ++d) shape[p][d] = rec.shape[p][d];
            const int k = subdivision_axes<D>(shape, st.iv, wk.node_level, order);
            if (k == 0) { at.add(&a.ctl->no_axis, 1ull); break; }              // getInverseOrder of an empty order raises (:1010)
            wk.nsplit = k;
#pragma unroll
            for (int p = 0; p < D; ++p)
#pragma unroll
                for (int j = 0; j < D; ++j) wk.order[p][j] = (j < k) ? order[p][j] : order[p][0];     // (value select: no run-time index)
            wk.node_index = -1;
            if (!wk.final_split) {
                const unsigned long long ni = at.add(&a.ctl->n_nodes, 1ull);
                if (ni < a.cap_nodes) {
                    NodeRec<D>& nd = a.nodes[ni];
                    for (int d = 0; d < D; ++d) { nd.iv[d][0] = wk.entry_iv[d][0]; nd.iv[d][1] = wk.entry_iv[d][1]; nd.next_mid[d] = st.next_mid[d]; }
                    nd.system = rec.system; nd.parent_node = rec.parent_node; nd.level = wk.node_level; nd.nsplit = k;
                    wk.node_index = (int)(ni + a.node_base);
                } else at.add(&a.ctl->overflow, 1ull);
            }
            int T, nmax[D]; unsigned long long total;
            box_sizes<D>(rec, wk, T, nmax, total);
            wk.op = kTSubdivide;
            dc.op = kTSubdivide; dc.need = need_subdivide(T, k); pick_team(dc.need, a.coarse, a.pol, 1, dc.cls); dc.bound = total << k;
            int e = 0; for (int d = 0; d < D; ++d) if (nmax[d] > e) e = nmax[d];
            dc.ext = (unsigned long long)e;
            dc.subdiv += 1;
        }
        break;
    }
}

// Level-0 boxes arrive packed from yr_init_boxes (tensors back to back, ld = n[D-1]): queue a copy-only pass.
template <int D, class At>
YR_HD void import_box(const Args<D>& a, unsigned int b, Decision& dc, At at) {
    const BoxRec<D>& rec = a.recs[b];
    Work<D>& w = a.work[b];
    unsigned long long off = rec.data_off;
    for (int p = 0; p < D; ++p) {
        w.off[p] = off;
        unsigned long long cnt = 1;
        for (int d = 0; d < D; ++d) { w.full[p][d] = rec.shape[p][d]; w.tshape[p][d] = rec.shape[p][d]; cnt *= (unsigned long long)rec.shape[p][d]; }
        off += cnt;
        w.ld[p] = rec.shape[p][D - 1];
        w.tsumabs[p] = 0.0; w.terr[p] = rec.err[p]; w.sumabs[p] = 0.0;
        for (int q = 0; q < kLow<D>; ++q) w.low[p][q] = 0.0;
        for (int j = 0; j < D; ++j) w.order[p][j] = 0;
    }
    for (int d = 0; d < D; ++d) {
        w.alpha[d] = 1.0; w.beta[d] = 0.0; w.last_size[d] = 0.0;
        w.entry_iv[d][0] = w.entry_iv[d][1] = 0.0;
        w.cell_iv[d][0] = rec.st.iv[d][0]; w.cell_iv[d][1] = rec.st.iv[d][1];
    }
    w.op = kTRestrict; w.fresh = 1; w.copy_only = 1; w.trim_valid = 0;
    w.zoom_count = 0; w.changed = 1; w.should_stop = 0; w.final_split = 0;
    w.node_level = rec.level; w.result_index = -1; w.node_index = -1; w.nsplit = 0;
    int T, nmax[D]; unsigned long long total;
    box_sizes<D>(rec, w, T, nmax, total);
    dc.op = kTRestrict; dc.need = need_restrict<D>(T, nmax); pick_team(dc.need, a.coarse, a.pol, 0, dc.cls); dc.bound = total;
    int e = 0; for (int d = 0; d < D; ++d) if (nmax[d] > e) e = nmax[d];
    dc.ext = (unsigned long long)e;
    at.add(&a.ctl->n_boxes, 1ull);
}

template <int D, class At>
YR_HD void commit_decision(const Args<D>& a, unsigned int b, const Decision& dc, At at) {
    if (dc.op == kTNone) return;
    const int slot = op_slot(dc.op);
    const unsigned long long pos = at.add(&a.ctl->n_list[slot][dc.cls], 1ull);
    a.lists[slot][dc.cls][pos] = b;
    at.max(&a.ctl->need_max[slot][dc.cls], dc.need);
    if (slot == 1) at.max(&a.ctl->ext_max[dc.cls], dc.ext);
    at.add(&a.ctl->out_bound, dc.bound);
}

}  // namespace fd
}  // namespace yr
