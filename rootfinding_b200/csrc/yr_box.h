// CTA-per-box pipeline of the Chebyshev subdivision solver.
//
// One cooperative thread array owns one frontier box from the moment its D
// coefficient tensors are staged into the workspace (shared memory, or a
// per-CTA global scratch for oversized boxes) until it is discarded, emitted
// as a root box, or replaced by its 2^k children written to the next level.
// Everything the reference does inside one solvePolyRecursive invocation --
// constant / quadratic checks, trimMs, the zoom loop (BoundingIntervalLinearSystem
// + interval restriction of every tensor along every axis), the final-step
// re-entry and the subdivision -- happens while the tensors stay on chip, so a
// box costs one read and its children one write of HBM.
//
// The code is written once against a tiny execution-context interface `Ctx`
// (thread id, barrier, block/warp sums, 64-bit atomics).  yr_kernels.cu binds
// it to CUDA; tests/emul binds it to pthreads so the same source can be checked
// against the oracle on a CPU-only machine.  The emulator is test code only.
//
// Reference (file:line in /root/reference/yroots/ChebyshevSubdivisionSolver.py):
//   build_matrices   :45-337   (TransformChebInPlace1D / ...ErrorFree / ...ErrorFreeSplit recurrences)
//   transform_axis   :45-120,:339-374 (the contraction; TransformChebInPlaceND)
//   trim_box         :1131-1171 (trimMs)
//   zoom loop        :896-956, :1243-1252 (zoomInOnIntervalIter, solvePolyRecursive)
//   subdivide_box    :981-1129 (getInverseOrder / getSubdivisionIntervals)
//   process_box      :1177-1317 (solvePolyRecursive without the tree-structured merging, which
//                    the host does on the small result set: rootfinding_b200/postpass.py)
#pragma once
#include "yr_types.h"

#ifndef YR_DEV
#if defined(__CUDACC__)
#define YR_DEV __device__
#define YR_DEVN __device__ __noinline__     /* large helpers exist once: the kernel is instruction-cache bound */
#else
#define YR_DEV
#define YR_DEVN
#endif
#endif

namespace yr {

constexpr int kStage = 16;      // register-staged outputs per thread for in-place restriction

enum Action : int {
    kActNone = 0, kActDiscard, kActEmit, kActReenterFinal, kActSubdivide, kActSubdivideFinal,
    kActZoomTransform, kActZoomStop
};

template <int D>
struct Layout {
    int base[D];        // tensors
    int tmp[D];         // subdivision temporaries (D-1 used)
    int cmat[2 * D];    // packed upper-triangular restriction matrices, column k at k(k+1)/2
    int rows_after;     // int[2D][nmax] stored in the double workspace
    int xscratch;       // exact mode: per matrix 6 columns of (nmax+2)
    int nmax, tri, max_size, total;
};

template <int D>
struct BoxShared {
    BoxRec<D> rec;
    Layout<D> lay;
    int shape[D][D];
    int stride[D][D];
    double err[D];
    double sumabs[D];
    double entry_iv[D][2];      // originalInterval.getIntervalForCombining() of the current invocation
    double cell_iv[D][2];
    double last_size[D];
    double alpha[2 * D], beta[2 * D];
    int mat_axis[2 * D], mat_n[2 * D];
    int nmat;
    int rows_buf[2][2 * D];
    double plane[64];
    double budget;
    int trim_stop, trimmed_any;
    int action;
    int state;
    int level;
    int zoom_count;
    int should_stop, changed;
    // subdivision
    int nsplit;
    int split_axes[D];          // ascending
    int order[D][D];
    int child_shape[1 << D][D][D];
    unsigned long long child_off[1 << D];
    unsigned long long child_rec_base;
    int node_index;
    int result_index;
    // per-CTA statistics, flushed once when the CTA retires
    unsigned long long c_boxes, c_zooms, c_subdiv, c_dconst, c_dquad, c_dzoom, c_coeffs, c_macs, c_maxlevel;
    unsigned long long work_index;
};

template <int D>
YR_HD int tensor_size(const int* shape) { int s = 1; for (int a = 0; a < D; ++a) s *= shape[a]; return s; }

// row-major decode of element e of a tensor view -> offset with the given strides
template <int D>
YR_HD int view_offset(int e, const int* shape, const int* stride) {
    int off = 0;
    for (int a = D - 1; a > 0; --a) { const int q = e / shape[a]; off += (e - q * shape[a]) * stride[a]; e = q; }
    return off + e * stride[0];
}

// fiber f (all axes except `axis`, row-major) -> offsets in source and destination
template <int D>
YR_HD void fiber_offsets(int f, int axis, const int* shape, const int* sstride, const int* dstride, int& so, int& dofs) {
    if (D == 1) { so = 0; dofs = 0; return; }
    if (D == 2) { so = f * sstride[1 - axis]; dofs = f * dstride[1 - axis]; return; }     // no integer division
    so = 0; dofs = 0;
    for (int a = D - 1; a >= 0; --a) {
        if (a == axis) continue;
        const int q = f / shape[a]; const int i = f - q * shape[a];
        so += i * sstride[a]; dofs += i * dstride[a]; f = q;
    }
}

// Rows of a tensor view = all index combinations of axes 0..D-2; the last axis is contiguous (stride 1
// survives every in-place restriction).  One division chain per row, none per element.
template <int D>
YR_HD int row_offset(int r, const int* shape, const int* stride) {
    int off = 0;
    for (int a = D - 2; a > 0; --a) { const int q = r / shape[a]; off += (r - q * shape[a]) * stride[a]; r = q; }
    return (D >= 2) ? off + r * stride[0] : 0;
}

template <int D, class Ctx>
YR_DEVN double tensor_abs_sum(Ctx& cx, const double* base, const int* shape, const int* stride) {
    int nrows = 1;
    for (int a = 0; a < D - 1; ++a) nrows *= shape[a];
    const int len = shape[D - 1];
    double s = 0;
    if (len >= 16 || D == 1) {            // lanes along the contiguous axis, warps over rows
        for (int r = cx.warp; r < nrows; r += cx.nwarps) {
            const double* row = base + row_offset<D>(r, shape, stride);
            for (int j = cx.lane; j < len; j += cx.lanes) s += fabs(row[j]);
        }
    } else {                              // short rows: threads over rows
        for (int r = cx.tid; r < nrows; r += cx.nthreads) {
            const double* row = base + row_offset<D>(r, shape, stride);
            for (int j = 0; j < len; ++j) s += fabs(row[j]);
        }
    }
    return cx.block_sum(s);
}

// Restriction matrices are stored row-major packed upper triangular: row i holds C[i][k], k = i..ncol-1.
YR_HD int crow_start(int i, int ncol) { return i * ncol - i * (i - 1) / 2; }

template <bool FMA>
YR_HD double tri_dot(const double* crow, int len, const double* s, int stride) {
    // sum_{j<len} crow[j] * s[j*stride], accumulated in ascending order like the reference (:84-112)
    double acc = 0.0;
    for (int j = 0; j < len; ++j) {
        if (FMA) acc = fma(crow[j], *s, acc); else acc = acc + crow[j] * *s;
        s += stride;
    }
    return acc;
}

// One fiber, all output rows, four rows per pass (they share the source loads; the matrix entries are
// the same for every lane of a warp, i.e. broadcast loads).  Safe in place: rows i..i+3 are stored after
// everything they read (source rows >= i) has been read, and later rows never read them.  Every row is
// accumulated in ascending source order starting from 0, exactly like tri_dot.
template <bool FMA>
YR_HDN double fiber_restrict(const double* C, int ncol, int n, int rows_out, const double* src, int sa, double* dst, int da) {
    double asum = 0.0;
    int i = 0;
    for (; i + 3 < rows_out; i += 4) {
        const double* c0 = C + crow_start(i, ncol);          // c0[j] = C[i][i+j]
        const double* c1 = C + crow_start(i + 1, ncol);      // c1[j] = C[i+1][i+1+j]
        const double* c2 = C + crow_start(i + 2, ncol);
        const double* c3 = C + crow_start(i + 3, ncol);
        const double* s = src + (long long)i * sa;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        {   // the triangular head: source rows i, i+1, i+2
            const double x0 = s[0], x1 = s[sa], x2 = s[2 * sa];
            if (FMA) {
                a0 = fma(c0[0], x0, a0); a0 = fma(c0[1], x1, a0); a0 = fma(c0[2], x2, a0);
                a1 = fma(c1[0], x1, a1); a1 = fma(c1[1], x2, a1);
                a2 = fma(c2[0], x2, a2);
            } else {
                a0 = a0 + c0[0] * x0; a0 = a0 + c0[1] * x1; a0 = a0 + c0[2] * x2;
                a1 = a1 + c1[0] * x1; a1 = a1 + c1[1] * x2;
                a2 = a2 + c2[0] * x2;
            }
        }
        s += 3 * sa;
        const int len = n - i - 3;                            // source rows i+3 .. n-1 feed all four
#pragma unroll 2
        for (int j = 0; j < len; ++j) {
            const double x = *s;
            if (FMA) { a0 = fma(c0[j + 3], x, a0); a1 = fma(c1[j + 2], x, a1); a2 = fma(c2[j + 1], x, a2); a3 = fma(c3[j], x, a3); }
            else { a0 = a0 + c0[j + 3] * x; a1 = a1 + c1[j + 2] * x; a2 = a2 + c2[j + 1] * x; a3 = a3 + c3[j] * x; }
            s += sa;
        }
        dst[(long long)i * da] = a0; dst[(long long)(i + 1) * da] = a1;
        dst[(long long)(i + 2) * da] = a2; dst[(long long)(i + 3) * da] = a3;
        asum += fabs(a0) + fabs(a1) + fabs(a2) + fabs(a3);
    }
    for (; i < rows_out; ++i) {
        const double v = tri_dot<FMA>(C + crow_start(i, ncol), n - i, src + (long long)i * sa, sa);
        dst[(long long)i * da] = v;
        asum += fabs(v);
    }
    return asum;
}

// dst = C applied along `axis` of src (rows_out output rows).  dst may be src (inplace).
// Returns sum|dst| to every thread.  Collective: all threads of the CTA must call it.
template <int D, class Ctx>
YR_DEVN double transform_axis(Ctx& cx, const double* src, const int* sshape, const int* sstride, int axis,
                             const double* C, int ncol, int rows_out, double* dst, const int* dstride,
                             bool inplace, bool use_fma) {
    const int n = sshape[axis];
    int F = 1;
    for (int a = 0; a < D; ++a) if (a != axis) F *= sshape[a];
    const int T = cx.nthreads;
    const int sa = sstride[axis], da = dstride[axis];
    double asum = 0;
    const long long items = (long long)rows_out * F;
    if (Ctx::kWarpPerBox || T <= 32 || F >= T || (inplace && rows_out > T * kStage)) {
        // a thread owns whole fibers: no barriers, no per-element index arithmetic
        for (int f = cx.tid; f < F; f += T) {
            int so, dofs; fiber_offsets<D>(f, axis, sshape, sstride, dstride, so, dofs);
            asum += use_fma ? fiber_restrict<true>(C, ncol, n, rows_out, src + so, sa, dst + dofs, da)
                            : fiber_restrict<false>(C, ncol, n, rows_out, src + so, sa, dst + dofs, da);
        }
    } else if (Ctx::kWarpPerBox) {
        // unreachable: lets the compiler drop the CTA-only paths from the warp-per-box kernel
    } else if (!inplace) {
        for (long long id = cx.tid; id < items; id += T) {
            const int i = (int)(id / F), f = (int)(id - (long long)i * F);
            int so, dofs; fiber_offsets<D>(f, axis, sshape, sstride, dstride, so, dofs);
            const double* crow = C + crow_start(i, ncol);
            const double* s0 = src + so + (long long)i * sa;
            const double v = use_fma ? tri_dot<true>(crow, n - i, s0, sa) : tri_dot<false>(crow, n - i, s0, sa);
            dst[dofs + (long long)i * da] = v;
            asum += fabs(v);
        }
    } else {
        // few fibers, many threads: spread (row, fiber) pairs over the CTA, stage results in registers,
        // barrier, then overwrite.  Fibers are taken in groups that fit the staging budget.
        int gf = (T * kStage) / rows_out; if (gf > F) gf = F; if (gf < 1) gf = 1;
        for (int f0 = 0; f0 < F; f0 += gf) {
            const int nf = (F - f0 < gf) ? (F - f0) : gf;
            const int cnt = rows_out * nf;
            double stage[kStage];
#pragma unroll
            for (int slot = 0; slot < kStage; ++slot) {
                const int id = cx.tid + slot * T;
                if (id < cnt) {
                    const int i = id / nf, f = f0 + (id - i * nf);
                    int so, dofs; fiber_offsets<D>(f, axis, sshape, sstride, dstride, so, dofs);
                    const double* crow = C + crow_start(i, ncol);
                    const double* s0 = src + so + (long long)i * sa;
                    stage[slot] = use_fma ? tri_dot<true>(crow, n - i, s0, sa) : tri_dot<false>(crow, n - i, s0, sa);
                }
            }
            cx.sync();
#pragma unroll
            for (int slot = 0; slot < kStage; ++slot) {
                const int id = cx.tid + slot * T;
                if (id < cnt) {
                    const int i = id / nf, f = f0 + (id - i * nf);
                    int so, dofs; fiber_offsets<D>(f, axis, sshape, sstride, dstride, so, dofs);
                    dst[dofs + (long long)i * da] = stage[slot];
                    asum += fabs(stage[slot]);
                }
            }
            cx.sync();
        }
    }
    return cx.block_sum(asum);
}

// --- restriction-matrix entries, exact (double-double) variants -------------------------------
struct ExactCols { double *p, *c, *q, *pe, *ce, *qe; };

YR_HD void exact_entry_general(const ExactCols& b, int i, int rows, double alpha, double beta,
                               dd as, dd bs, double& V, double& E) {
    // TransformChebInPlace1DErrorFree :166-227
    if (i == 0) {
        dd v1 = two_prod_split(beta, 2 * b.c[0], bs.hi, bs.lo);
        dd v2 = two_prod_split(alpha, b.c[1], as.hi, as.lo);
        dd v3 = two_sum(v1.hi, v2.hi);
        dd v4 = two_sum(v3.hi, -b.p[0]);
        V = v4.hi;
        E = -b.pe[0] + alpha * b.ce[1] + 2 * beta * b.ce[0] + v1.lo + v2.lo + v3.lo + v4.lo;
    } else if (i == rows) {
        dd f = two_prod_split(alpha, b.c[rows - 1], as.hi, as.lo);
        E = f.lo + alpha * b.ce[rows - 1];
        V = f.hi;
    } else if (i == rows - 1) {
        const double c1 = (i == 1) ? 2.0 : 1.0;
        dd v1 = two_prod_split(beta, 2 * b.c[i], bs.hi, bs.lo);
        dd v2 = two_prod_split(alpha, c1 * b.c[i - 1], as.hi, as.lo);
        dd v3 = two_sum(v1.hi, v2.hi);
        dd v4 = two_sum(v3.hi, -b.p[i]);
        V = v4.hi;
        E = -b.pe[i] + c1 * alpha * b.ce[i - 1] + 2 * beta * b.ce[i] + v1.lo + v2.lo + v3.lo + v4.lo;
    } else if (i == 1) {
        dd v1 = two_sum(2 * b.c[0], b.c[2]);
        dd v2 = two_prod_split(beta, 2 * b.c[1], bs.hi, bs.lo);
        dd v3 = two_prod_split(alpha, v1.hi, as.hi, as.lo);
        dd v4 = two_sum(v2.hi, v3.hi);
        dd v5 = two_sum(v4.hi, -b.p[1]);
        V = v5.hi;
        E = -b.pe[1] + alpha * (2 * b.ce[0] + b.ce[2] + v1.lo) + 2 * beta * b.ce[1] + v2.lo + v3.lo + v4.lo + v5.lo;
    } else {
        dd v1 = two_sum(b.c[i - 1], b.c[i + 1]);
        dd v2 = two_prod_split(beta, 2 * b.c[i], bs.hi, bs.lo);
        dd v3 = two_prod_split(alpha, v1.hi, as.hi, as.lo);
        dd v4 = two_sum(v2.hi, v3.hi);
        dd v5 = two_sum(v4.hi, -b.p[i]);
        V = v5.hi;
        E = -b.pe[i] + alpha * (b.ce[i - 1] + b.ce[i + 1] + v1.lo) + 2 * beta * b.ce[i] + v2.lo + v3.lo + v4.lo + v5.lo;
    }
}

YR_HD void exact_entry_split(const ExactCols& b, int i, int rows, double sgn, double& V, double& E) {
    // TransformChebInPlace1DErrorFreeSplit :278-326
    if (i == 0) {
        dd v1 = two_sum(b.c[1] / 2, sgn * b.c[0]);
        dd v2 = two_sum(v1.hi, -b.p[0]);
        V = v2.hi;
        E = -b.pe[0] + b.ce[1] / 2 + sgn * b.ce[0] + v1.lo + v2.lo;
    } else if (i == rows) {
        V = b.c[rows - 1] / 2;
        E = b.ce[rows - 1] / 2;
    } else if (i == rows - 1) {
        const double c1 = (i == 1) ? 1.0 : 0.5;
        dd v1 = two_sum(c1 * b.c[i - 1], sgn * b.c[i]);
        dd v2 = two_sum(v1.hi, -b.p[i]);
        V = v2.hi;
        E = -b.pe[i] + c1 * b.ce[i - 1] + sgn * b.ce[i] + v1.lo + v2.lo;
    } else if (i == 1) {
        dd v1 = two_sum(b.c[0], b.c[2] / 2);
        dd v2 = two_sum(v1.hi, sgn * b.c[1]);
        dd v3 = two_sum(v2.hi, -b.p[1]);
        V = v3.hi;
        E = -b.pe[1] + b.ce[0] + b.ce[2] / 2 + sgn * b.ce[1] + v1.lo + v2.lo + v3.lo;
    } else {
        dd v1 = two_sum(b.c[i - 1], b.c[i + 1]);
        dd v2 = two_sum(v1.hi / 2, sgn * b.c[i]);
        dd v3 = two_sum(v2.hi, -b.p[i]);
        V = v3.hi;
        E = -b.pe[i] + (b.ce[i - 1] + b.ce[i + 1] + v1.lo) / 2 + sgn * b.ce[i] + v2.lo + v3.lo;
    }
}

template <int D, class Ctx>
YR_DEVN void build_matrices_warp(Ctx& cx, BoxShared<D>& sh, double* ws);

// Builds sh.nmat restriction matrices C(alpha_m, beta_m) with mat_n[m] columns each, column by
// column (the three-term recurrence couples neighbouring rows, so columns are sequential and rows
// parallel), together with rows_after[m][k] = the reference's `maxRow` after column k.
template <int D, class Ctx>
YR_DEVN void build_matrices(Ctx& cx, BoxShared<D>& sh, double* ws, bool exact) {
    const Layout<D>& L = sh.lay;
    int* rows_after = reinterpret_cast<int*>(ws + L.rows_after);
    const int nmat = sh.nmat;
    int kmax = 0;
    for (int m = 0; m < nmat; ++m) kmax = sh.mat_n[m] > kmax ? sh.mat_n[m] : kmax;
    if constexpr (Ctx::kWarpPerBox) {
        if (!exact && kmax <= cx.lanes) { build_matrices_warp<D>(cx, sh, ws); return; }
    }
    const int xs = L.nmax + 2;
    for (int m = cx.tid; m < nmat; m += cx.nthreads) {
        double* C = ws + L.cmat[m];
        const int nc = sh.mat_n[m];
        C[0] = 1.0;                                              // C[0][0]
        rows_after[m * L.nmax + 0] = 1;
        sh.rows_buf[0][m] = 1;
        if (nc >= 2) {
            C[1] = sh.beta[m]; C[crow_start(1, nc)] = sh.alpha[m];   // C[0][1], C[1][1]
            rows_after[m * L.nmax + 1] = 2;
        }
        sh.rows_buf[1][m] = 2;
    }
    if (exact) {
        for (int id = cx.tid; id < nmat * 6 * xs; id += cx.nthreads) ws[L.xscratch + id] = 0.0;
        cx.sync();
        for (int m = cx.tid; m < nmat; m += cx.nthreads) {
            double* X = ws + L.xscratch + m * 6 * xs;
            const bool split = (sh.alpha[m] == 0.5 && fabs(sh.beta[m]) == 0.5);
            X[0 * xs + 0] = 1.0;                                   // column 0 values
            X[1 * xs + 0] = split ? (sh.beta[m] > 0 ? 0.5 : -0.5) : sh.beta[m];
            X[1 * xs + 1] = split ? 0.5 : sh.alpha[m];
        }
    }
    cx.sync();
    for (int k = 2; k < kmax; ++k) {
        for (int m = 0; m < nmat; ++m)
        for (int i = cx.tid; i <= k; i += cx.nthreads) {
            if (k >= sh.mat_n[m]) continue;
            const int nc = sh.mat_n[m];
            const int rows = sh.rows_buf[(k - 1) & 1][m];
            double* C = ws + L.cmat[m];
            const double alpha = sh.alpha[m], beta = sh.beta[m];
            double val = 0.0;
            if (!exact) {
                if (i <= rows) {
                    // entries beyond a column's length are zero, which reproduces the reference's edge formulas exactly
                    const double cm1 = (i >= 1) ? C[crow_start(i - 1, nc) + (k - 1) - (i - 1)] : 0.0;
                    const double ci = (i <= k - 1) ? C[crow_start(i, nc) + (k - 1) - i] : 0.0;
                    const double cp1 = (i + 1 <= k - 1) ? C[crow_start(i + 1, nc) + (k - 1) - (i + 1)] : 0.0;
                    const double pi = (i <= k - 2) ? C[crow_start(i, nc) + (k - 2) - i] : 0.0;
                    double t;
                    if (i == 0) t = alpha * cp1;
                    else if (i == 1) t = alpha * (2 * cm1 + cp1);
                    else t = alpha * (cm1 + cp1);
                    val = (-pi + t) + (2 * beta) * ci;
                    if (i == rows) {                                // the reference's finalVal, :106-112
                        val = alpha * cm1;
                        int nr = rows;
                        if (fabs(val) > 1e-16) nr = rows + 1; else val = 0.0;
                        sh.rows_buf[k & 1][m] = nr;
                        rows_after[m * L.nmax + k] = nr;
                    }
                }
            } else {
                double* X = ws + L.xscratch + m * 6 * xs;
                ExactCols b;
                b.p = X + ((k - 2) % 3) * xs; b.c = X + ((k - 1) % 3) * xs; b.q = X + (k % 3) * xs;
                b.pe = b.p + 3 * xs; b.ce = b.c + 3 * xs; b.qe = b.q + 3 * xs;
                if (i <= rows) {
                    double V, E;
                    const bool split = (alpha == 0.5 && fabs(beta) == 0.5);
                    if (split) exact_entry_split(b, i, rows, beta > 0 ? 1.0 : -1.0, V, E);
                    else exact_entry_general(b, i, rows, alpha, beta, veltkamp(alpha), veltkamp(beta), V, E);
                    b.q[i] = V; b.qe[i] = E;
                    val = V + E;
                    if (i == rows) {                                // always accumulated, grows above 1e-32 (:222-227)
                        const int nr = (fabs(val) > 1e-32) ? rows + 1 : rows;
                        sh.rows_buf[k & 1][m] = nr;
                        rows_after[m * L.nmax + k] = nr;
                    }
                }
            }
            C[crow_start(i, nc) + k - i] = val;
        }
        cx.sync();
    }
}

// Warp-per-box variant of build_matrices (plain arithmetic, extents <= warp width): lane i keeps row i of
// the two live columns in registers, neighbours come from shuffles, no shared-memory reads and no barriers.
// Same expressions, same order, same truncation rule -> bit-identical matrices.
template <int D, class Ctx>
YR_DEVN void build_matrices_warp(Ctx& cx, BoxShared<D>& sh, double* ws) {
    const Layout<D>& L = sh.lay;
    int* rows_after = reinterpret_cast<int*>(ws + L.rows_after);
    const int i = cx.lane;
    for (int m = 0; m < sh.nmat; ++m) {
        const int nc = sh.mat_n[m];
        double* C = ws + L.cmat[m];
        const double alpha = sh.alpha[m], beta = sh.beta[m];
        double* row = C + crow_start(i, nc) - i;                 // row[k] == C[i][k] for k >= i
        double prev2 = (i == 0) ? 1.0 : 0.0;                      // column 0
        double prev = (i == 0) ? beta : (i == 1 ? alpha : 0.0);   // column 1
        if (i == 0) { row[0] = 1.0; rows_after[m * L.nmax] = 1; if (nc >= 2) { row[1] = beta; rows_after[m * L.nmax + 1] = 2; } }
        if (i == 1 && nc >= 2) row[1] = alpha;
        int rows = 2;
        for (int k = 2; k < nc; ++k) {
            double cm1 = cx.shfl_up(prev, 1), cp1 = cx.shfl_down(prev, 1);
            if (i == 0) cm1 = 0.0;
            if (i + 1 > k - 1) cp1 = 0.0;
            double val = 0.0;
            int grow = 0;
            if (i <= rows) {
                double t;
                if (i == 0) t = alpha * cp1;
                else if (i == 1) t = alpha * (2 * cm1 + cp1);
                else t = alpha * (cm1 + cp1);
                val = (-prev2 + t) + (2 * beta) * prev;
                if (i == rows) {                                  // the reference's finalVal, :106-112
                    val = alpha * cm1;
                    if (fabs(val) > 1e-16) grow = 1; else val = 0.0;
                }
            }
            grow = cx.shfl_idx_int(grow, rows);
            if (i <= k) row[k] = val;
            prev2 = prev; prev = val;
            rows += grow;
            if (i == 0) rows_after[m * L.nmax + k] = rows;
        }
    }
    cx.sync();
}

// rows the reference returns for a tensor of extent n along the axis of matrix m
template <int D>
YR_HD int rows_for(const BoxShared<D>& sh, const double* ws, int m, int n) {
    const int* rows_after = reinterpret_cast<const int*>(ws + sh.lay.rows_after);
    return rows_after[m * sh.lay.nmax + (n - 1)];
}

template <int D>
YR_HD void compute_layout(BoxShared<D>& sh, bool exact, bool subdivide = true) {
    Layout<D>& L = sh.lay;
    int off = 0, mx = 0, nmax = 1;
    for (int p = 0; p < D; ++p) {
        const int sz = tensor_size<D>(sh.shape[p]);
        L.base[p] = off; off += sz; mx = sz > mx ? sz : mx;
        for (int a = 0; a < D; ++a) nmax = sh.shape[p][a] > nmax ? sh.shape[p][a] : nmax;
    }
    L.total = off; L.max_size = mx; L.nmax = nmax; L.tri = nmax * (nmax + 1) / 2;
    const int nmat = subdivide ? 2 * D : D;
    for (int j = 0; j < D; ++j) { L.tmp[j] = off; if (subdivide && j < D - 1) off += mx; }
    for (int m = 0; m < 2 * D; ++m) { L.cmat[m] = off; if (m < nmat) off += L.tri; }
    L.rows_after = off; off += (2 * D * nmax + 1) / 2 + 1;
    L.xscratch = off;
    if (exact) off += nmat * 6 * (nmax + 2);
}

// ---------------------------------------------------------------------------------------------
// The alternative trim policies of the reference's experiment file (tests/TrimMs optimization/
// ChebyshevSubdivisionSolverWithOptimizedTrimMs.py, "T:"): 1 = trimMsOptimized1 T:1167-1197 (one hyperplane per axis per
// sweep), 2 = trimMsOptimized2 T:1199-1235 (always the axis with the smallest last hyperplane), 3 = trimMsOptimized3
// T:1237-1271 (axes by descending extent, each as far as it goes; the running total is added to the error after EVERY
// axis, as T:1271 does).  Extents stop at 2 (trimMs: 3).  Every thread carries the same extents / budget / error, the
// plane sums are team-wide reductions.
template <int D, class Ctx>
YR_DEVN void trim_box_policy(Ctx& cx, BoxShared<D>& sh, double* ws, const SolverParams& prm) {
    for (int p = 0; p < D; ++p) {
        const double* base = ws + sh.lay.base[p];
        int shape[D], stride[D];
        for (int a = 0; a < D; ++a) { shape[a] = sh.shape[p][a]; stride[a] = sh.stride[p][a]; }
        double err = sh.err[p];
        double budget = prm.trim_abs_tol + err * prm.trim_rel_tol;
        bool any = false;
        auto last_sum = [&](int axis) {
            int shp[D];
            for (int a = 0; a < D; ++a) shp[a] = shape[a];
            shp[axis] = 1;
            return tensor_abs_sum<D>(cx, base + (long long)(shape[axis] - 1) * stride[axis], shp, stride);
        };
        if (prm.trim_policy == 1) {
            bool again = true;
            while (again) {
                again = false;
                double swept = 0;
                for (int axis = 0; axis < D; ++axis) {
                    const double s = last_sum(axis);
                    if (shape[axis] > 2 && s < budget) { again = true; shape[axis] -= 1; budget -= s; swept += s; any = true; }
                }
                err += swept;
            }
        } else if (prm.trim_policy == 2) {
            for (;;) {
                double best = INFINITY; int best_axis = -1;
                for (int axis = 0; axis < D; ++axis) {
                    const double s = last_sum(axis);
                    if (s < best && shape[axis] > 2) { best = s; best_axis = axis; }
                }
                if (best_axis < 0 || best >= budget) break;
                shape[best_axis] -= 1; budget -= best; err += best; any = true;
            }
        } else {
            // np.argsort(shape)[::-1]: ascending stable sort of the extents, reversed
            int order[D];
            for (int a = 0; a < D; ++a) order[a] = a;
            for (int a = 1; a < D; ++a) { const int v = order[a]; int q = a - 1; while (q >= 0 && shape[order[q]] > shape[v]) { order[q + 1] = order[q]; --q; } order[q + 1] = v; }
            double total = 0;
            for (int j = D - 1; j >= 0; --j) {
                const int axis = order[j];
                for (;;) {
                    const double s = last_sum(axis);
                    if (s < budget && shape[axis] > 2) { shape[axis] -= 1; budget -= s; total += s; any = true; }
                    else break;
                }
                err += total;
            }
        }
        cx.sync();
        if (cx.tid == 0) {
            for (int a = 0; a < D; ++a) sh.shape[p][a] = shape[a];
            sh.err[p] = err;
            if (any) sh.trimmed_any = 1;
        }
        cx.sync();
    }
}

template <int D, class Ctx>
YR_DEVN void trim_box(Ctx& cx, BoxShared<D>& sh, double* ws, const SolverParams& prm) {
    if (prm.trim_policy != 0) { trim_box_policy<D>(cx, sh, ws, prm); return; }
    for (int p = 0; p < D; ++p) {
        if (cx.tid == 0) sh.budget = prm.trim_abs_tol + sh.err[p] * prm.trim_rel_tol;
        cx.sync();
        const double* base = ws + sh.lay.base[p];
        for (int axis = 0; axis < D; ++axis) {
            for (;;) {
                const int n = sh.shape[p][axis];
                if (n <= 3) break;
                int nb = n - 3; if (nb > cx.nwarps) nb = cx.nwarps; if (nb > 64) nb = 64;
                // warp w sums the plane of index n-1-w (current extents of the other axes)
                double s = 0;
                if (cx.warp < nb) {
                    int shp[D];
                    for (int a = 0; a < D; ++a) shp[a] = sh.shape[p][a];
                    shp[axis] = 1;
                    const double* plane = base + (long long)(n - 1 - cx.warp) * sh.stride[p][axis];
                    int nrows = 1;
                    for (int a = 0; a < D - 1; ++a) nrows *= shp[a];
                    const int len = shp[D - 1];
                    if (len >= 16 || D == 1) {
                        for (int r = 0; r < nrows; ++r) {
                            const double* row = plane + row_offset<D>(r, shp, sh.stride[p]);
                            for (int j = cx.lane; j < len; j += cx.lanes) s += fabs(row[j]);
                        }
                    } else {
                        for (int r = cx.lane; r < nrows; r += cx.lanes) {
                            const double* row = plane + row_offset<D>(r, shp, sh.stride[p]);
                            for (int j = 0; j < len; ++j) s += fabs(row[j]);
                        }
                    }
                }
                s = cx.warp_sum(s);
                if (cx.warp < nb && cx.lane == 0) sh.plane[cx.warp] = s;
                cx.sync();
                if (cx.tid == 0) {
                    sh.trim_stop = 0;
                    for (int w = 0; w < nb; ++w) {
                        if (sh.plane[w] < sh.budget) {
                            sh.shape[p][axis] -= 1;
                            sh.budget -= sh.plane[w];
                            sh.err[p] += sh.plane[w];
                            sh.trimmed_any = 1;
                        } else { sh.trim_stop = 1; break; }
                    }
                }
                cx.sync();
                if (sh.trim_stop) break;
            }
        }
    }
}

template <int D>
YR_HD void fill_result(const BoxShared<D>& sh, ResultRec<D>& r, uint32_t extra_flags) {
    const IntervalState<D>& st = sh.rec.st;
    const bool fin = st.final_step();
    for (int d = 0; d < D; ++d) {
        // reported box: the map frozen at startFinalStep (getFinalInterval :460-489)
        const dd A = fin ? st.pfA[d] : st.A[d], B = fin ? st.pfB[d] : st.B[d];
        dd nA; nA.hi = -A.hi; nA.lo = -A.lo;
        r.box[d][0] = dd_to_double(dd_add(B, nA));
        r.box[d][1] = dd_to_double(dd_add(B, A));
        // reported point: midpoint of the current (post final-step) interval (getFinalPoint :491-514)
        dd nA2; nA2.hi = -st.A[d].hi; nA2.lo = -st.A[d].lo;
        const double lo = dd_to_double(dd_add(st.B[d], nA2));
        const double hi = dd_to_double(dd_add(st.B[d], st.A[d]));
        r.root[d] = fin ? (lo + hi) / 2 : (r.box[d][0] + r.box[d][1]) / 2;
        r.comb_iv[d][0] = fin ? st.pf_iv[d][0] : st.iv[d][0];
        r.comb_iv[d][1] = fin ? st.pf_iv[d][1] : st.iv[d][1];
        r.iv_lo[d] = st.iv[d][0];
        r.cell_iv[d][0] = sh.cell_iv[d][0]; r.cell_iv[d][1] = sh.cell_iv[d][1];
        r.next_mid[d] = st.next_mid[d];
    }
    bool touch = false;
    for (int d = 0; d < D; ++d)
        touch = touch || r.comb_iv[d][0] == sh.cell_iv[d][0] || r.comb_iv[d][1] == sh.cell_iv[d][1];
    r.system = sh.rec.system; r.parent_node = sh.rec.parent_node; r.collector = sh.rec.collector;
    r.level = sh.level;
    r.flags = extra_flags | (fin ? kResFinalStep : 0u) | (touch ? kResTouchesCell : 0u);
    r.pad = 0;
    r.path_hi = sh.rec.path_hi; r.path_lo = sh.rec.path_lo;
}

template <int D, class Ctx>
YR_DEVN void emit_result(Ctx& cx, const LevelArgs<D>& args, BoxShared<D>& sh, uint32_t extra_flags) {
    if (cx.tid == 0) {
        const unsigned long long idx = cx.atomic_add(&args.ctr->n_results, 1ull);
        sh.result_index = (int)(idx + args.result_base);
        if (idx < args.cap_results) fill_result<D>(sh, args.results[idx], extra_flags);
        else cx.atomic_add(&args.ctr->overflow, 1ull);
    }
}

// Scalar gather of the low-order coefficients the checks need (thread 0 only).
template <int D>
YR_HD double coef_at(const BoxShared<D>& sh, const double* ws, int p, const int* idx) {
    int off = 0;
    for (int a = 0; a < D; ++a) { if (idx[a] >= sh.shape[p][a]) return 0.0; off += idx[a] * sh.stride[p][a]; }
    return ws[sh.lay.base[p] + off];
}

template <int D>
YR_HDN bool quadratic_check_box(const BoxShared<D>& sh, const double* ws, int p, double tol) {
    int idx[D];
    auto at = [&](int a, int va, int b, int vb) {
        for (int d = 0; d < D; ++d) idx[d] = 0;
        if (a >= 0) idx[a] = va;
        if (b >= 0) idx[b] = vb;
        return coef_at<D>(sh, ws, p, idx);
    };
    if (D == 2) {
        const double c[6] = {at(-1, 0, -1, 0), at(0, 1, -1, 0), at(1, 1, -1, 0), at(0, 2, -1, 0), at(0, 1, 1, 1), at(1, 2, -1, 0)};
        return quad_check_2d(c, sh.sumabs[p], tol);
    } else if (D == 3) {
        const double c[10] = {at(-1, 0, -1, 0), at(0, 1, -1, 0), at(1, 1, -1, 0), at(2, 1, -1, 0), at(0, 1, 1, 1),
                              at(0, 1, 2, 1), at(1, 1, 2, 1), at(0, 2, -1, 0), at(1, 2, -1, 0), at(2, 2, -1, 0)};
        return quad_check_3d(c, sh.sumabs[p], tol);
    } else {
        double g[D], pure[D], cross[D][D];
        const double c0 = at(-1, 0, -1, 0);
        double used = fabs(c0);
        for (int a = 0; a < D; ++a) { g[a] = at(a, 1, -1, 0); pure[a] = at(a, 2, -1, 0); used += fabs(g[a]) + fabs(pure[a]); }
        for (int a = 0; a < D; ++a) for (int b = 0; b < D; ++b) cross[a][b] = 0.0;
        for (int a = 0; a < D; ++a) for (int b = a + 1; b < D; ++b) { cross[a][b] = at(a, 1, b, 1); used += fabs(cross[a][b]); }
        return quad_check_nd<D>(c0, g, pure, cross, sh.sumabs[p] - used, tol);
    }
}

// ---------------------------------------------------------------------------------------------
// Subdivision: 2^k children written straight to the next level's pool.
template <int D, class Ctx>
YR_DEVN void subdivide_box(Ctx& cx, const LevelArgs<D>& args, BoxShared<D>& sh, double* ws, bool final_step_split) {
    const SolverParams& prm = args.prm;
    if (cx.tid == 0) {
        const int k = subdivision_axes<D>(sh.shape, sh.rec.st.iv, sh.level, sh.order);
        sh.nsplit = k;
        if (k == 0) cx.atomic_add(&args.ctr->no_axis, 1ull);          // the reference raises here (:1010); the host reports it
        // ascending set of split axes; matrix 2j (lower half) and 2j+1 (upper half) of axis j
        bool used[D];
        for (int a = 0; a < D; ++a) used[a] = false;
        for (int j = 0; j < k; ++j) used[sh.order[0][j]] = true;
        int j = 0;
        for (int a = 0; a < D; ++a) if (used[a]) sh.split_axes[j++] = a;
        for (j = 0; j < k; ++j) {
            const int a = sh.split_axes[j];
            const double mid = sh.rec.st.next_mid[a];
            const double alpha = (mid + 1) / 2, beta = (mid - 1) / 2;       // :1089
            int n = 1;
            for (int p = 0; p < D; ++p) n = sh.shape[p][a] > n ? sh.shape[p][a] : n;
            sh.alpha[2 * j] = alpha;  sh.beta[2 * j] = beta;   sh.mat_axis[2 * j] = a;     sh.mat_n[2 * j] = n;
            sh.alpha[2 * j + 1] = -beta; sh.beta[2 * j + 1] = alpha; sh.mat_axis[2 * j + 1] = a; sh.mat_n[2 * j + 1] = n;
        }
        sh.nmat = 2 * k;
    }
    cx.sync();
    build_matrices<D>(cx, sh, ws, prm.exact != 0);
    const int k = sh.nsplit;
    const int nchild = 1 << k;
    if (cx.tid == 0) {
        unsigned long long total = 0, biggest = 0, widest = 0, fattest = 0;
        for (int c = 0; c < nchild; ++c) {
            unsigned long long csz = 0;
            sh.child_off[c] = total;
            for (int p = 0; p < D; ++p) {
                for (int a = 0; a < D; ++a) sh.child_shape[c][p][a] = sh.shape[p][a];
                for (int j = 0; j < k; ++j) {
                    const int a = sh.split_axes[j];
                    const int half = (c >> (k - 1 - j)) & 1;
                    if (sh.shape[p][a] > 1) sh.child_shape[c][p][a] = rows_for<D>(sh, ws, 2 * j + half, sh.shape[p][a]);
                }
                const unsigned long long tsz = (unsigned long long)tensor_size<D>(sh.child_shape[c][p]);
                csz += tsz; fattest = tsz > fattest ? tsz : fattest;
                for (int a = 0; a < D; ++a) widest = (unsigned long long)sh.child_shape[c][p][a] > widest ? sh.child_shape[c][p][a] : widest;
            }
            total += (csz + 1ull) & ~1ull;      // every box starts 16-byte aligned (TMA bulk copies)
            biggest = csz > biggest ? csz : biggest;
        }
        const unsigned long long rb = cx.atomic_add(&args.ctr->n_children, (unsigned long long)nchild);
        const unsigned long long db = cx.atomic_add(&args.ctr->data_used, total);
        cx.atomic_max(&args.ctr->max_child_doubles, biggest);
        cx.atomic_max(&args.ctr->max_child_extent, widest);
        cx.atomic_max(&args.ctr->max_child_tensor, fattest);
        sh.child_rec_base = rb;
        bool ok = (rb + nchild <= args.cap_recs_out) && (db + total <= args.cap_data_out);
        sh.node_index = -1;
        if (!final_step_split) {
            const unsigned long long ni = cx.atomic_add(&args.ctr->n_nodes, 1ull);
            if (ni < args.cap_nodes) {
                NodeRec<D>& nd = args.nodes[ni];
                for (int d = 0; d < D; ++d) { nd.iv[d][0] = sh.entry_iv[d][0]; nd.iv[d][1] = sh.entry_iv[d][1]; nd.next_mid[d] = sh.rec.st.next_mid[d]; }
                nd.system = sh.rec.system; nd.parent_node = sh.rec.parent_node; nd.level = sh.level; nd.nsplit = k;
                sh.node_index = (int)(ni + args.node_base);
            } else ok = false;
        }
        if (!ok) { cx.atomic_add(&args.ctr->overflow, 1ull); sh.action = kActDiscard; }
        for (int c = 0; c < nchild; ++c) sh.child_off[c] += db;
        sh.c_subdiv += 1;
    }
    cx.sync();
    if (sh.action == kActDiscard) return;      // overflow: the host re-runs the level with larger buffers
    // child records (thread c writes child c)
    for (int c = cx.tid; c < nchild; c += cx.nthreads) {
        BoxRec<D>& out = args.recs_out[sh.child_rec_base + c];
        out.data_off = sh.child_off[c];
        for (int p = 0; p < D; ++p) { for (int a = 0; a < D; ++a) out.shape[p][a] = sh.child_shape[c][p][a]; out.err[p] = 0.0; }
        IntervalState<D> st = sh.rec.st;
        for (int j = 0; j < k; ++j) {                                  // :1111-1128
            const int a = sh.split_axes[j];
            const int half = (c >> (k - 1 - j)) & 1;
            const double mid = st.next_mid[a];
            double sub[D][2], al[D], be[D];
            for (int d = 0; d < D; ++d) { sub[d][0] = -1.0; sub[d][1] = 1.0; }
            if (half) sub[a][0] = mid; else sub[a][1] = mid;
            add_transform<D>(st, sub, al, be);
        }
        for (int j = 0; j < k; ++j) st.next_mid[sh.split_axes[j]] = 0.0;
        out.st = st;
        out.system = sh.rec.system;
        out.parent_node = final_step_split ? sh.rec.parent_node : sh.node_index;
        out.collector = final_step_split ? sh.result_index : sh.rec.collector;
        out.level = sh.level;
        // append the k-bit child slot below the bits already used (left aligned 128-bit key)
        unsigned long long hi = sh.rec.path_hi, lo = sh.rec.path_lo;
        const int used = sh.rec.path_bits;
        const int shift = 128 - used - k;
        if (shift >= 64) hi |= (unsigned long long)c << (shift - 64);
        else if (shift >= 0) {
            lo |= (unsigned long long)c << shift;
            if (shift + k > 64) hi |= (unsigned long long)c >> (64 - shift);
        }
        out.path_bits = (shift >= 0) ? used + k : used;
        out.pad0 = 0;
        out.path_hi = hi; out.path_lo = lo;
    }
    // tensors: per polynomial, halve along its own axis order (largest extent first), depth first,
    // keeping the intermediate pieces in the workspace and writing the last stage to HBM.
    for (int p = 0; p < D; ++p) {
        const double* M = ws + sh.lay.base[p];
        int tshape[D + 1][D], tstride[D + 1][D];
        double tsum[D + 1], terr[D + 1];
        for (int a = 0; a < D; ++a) { tshape[0][a] = sh.shape[p][a]; tstride[0][a] = sh.stride[p][a]; }
        tsum[0] = sh.sumabs[p]; terr[0] = sh.err[p];
        int rank[D];                                                  // position of an axis in the ascending set
        for (int j = 0; j < k; ++j) rank[sh.split_axes[j]] = j;
        for (int combo = 0; combo < nchild; ++combo) {
            // first stage whose choice differs from the previous combo
            int first = 0;
            if (combo > 0) {
                const int diff = combo ^ (combo - 1);
                for (first = 0; first < k; ++first) if ((diff >> (k - 1 - first)) & 1) break;
            }
            int canon = 0;
            for (int s = 0; s < k; ++s) canon |= ((combo >> (k - 1 - s)) & 1) << (k - 1 - rank[sh.order[p][s]]);
            for (int s = first; s < k; ++s) {
                const int axis = sh.order[p][s];
                const int half = (combo >> (k - 1 - s)) & 1;
                const int m = 2 * rank[axis] + half;
                const double* src = (s == 0) ? M : ws + sh.lay.tmp[s - 1];
                const int n = tshape[s][axis];
                const double e1 = n * kMachEps * tsum[s];             // getTransformationError on the piece being halved
                terr[s + 1] = e1 + terr[s];
                const bool last = (s == k - 1);
                double* dst = last ? args.data_out + sh.child_off[canon] : ws + sh.lay.tmp[s];
                if (last) for (int q = 0; q < p; ++q) dst += tensor_size<D>(sh.child_shape[canon][q]);
                for (int a = 0; a < D; ++a) tshape[s + 1][a] = tshape[s][a];
                int rows = n;
                if (n > 1) rows = rows_for<D>(sh, ws, m, n);
                tshape[s + 1][axis] = rows;
                int st = 1;
                for (int a = D - 1; a >= 0; --a) { tstride[s + 1][a] = st; st *= tshape[s + 1][a]; }
                if (n > 1) {
                    tsum[s + 1] = transform_axis<D>(cx, src, tshape[s], tstride[s], axis, ws + sh.lay.cmat[m], sh.mat_n[m], rows,
                                                    dst, tstride[s + 1], false, prm.use_fma != 0);
                    if (cx.tid == 0) {
                        unsigned long long fibers = 1;
                        for (int a = 0; a < D; ++a) if (a != axis) fibers *= (unsigned long long)tshape[s][a];
                        sh.c_macs += fibers * ((unsigned long long)rows * n - (unsigned long long)rows * (rows - 1) / 2);
                    }
                } else {
                    // extent-1 axis: TransformChebInPlaceND returns the tensor unchanged (:363)
                    const int cnt = tensor_size<D>(tshape[s]);
                    for (int e = cx.tid; e < cnt; e += cx.nthreads)
                        dst[view_offset<D>(e, tshape[s + 1], tstride[s + 1])] = src[view_offset<D>(e, tshape[s], tstride[s])];
                    tsum[s + 1] = tsum[s];
                    cx.sync();
                }
                if (!last) cx.sync();
            }
            if (cx.tid == 0) args.recs_out[sh.child_rec_base + canon].err[p] = terr[k];
        }
    }
}

// transformChebToInterval :864-894 on the staged box: sh.alpha/beta[axis] hold the restriction,
// every tensor is restricted in place along every axis, errors are charged before each axis.
template <int D, class Ctx>
YR_DEVN void restrict_system_inplace(Ctx& cx, BoxShared<D>& sh, double* ws, const SolverParams& prm) {
    build_matrices<D>(cx, sh, ws, prm.exact != 0);
    for (int p = 0; p < D; ++p) {
        double s = sh.sumabs[p];
        for (int axis = 0; axis < D; ++axis) {
            const int n = sh.shape[p][axis];
            if (cx.tid == 0) sh.err[p] += n * kMachEps * s;            // charged before the axis (:860)
            const bool identity = (sh.alpha[axis] == 1.0 && sh.beta[axis] == 0.0) || n == 1;
            if (!identity) {
                const int rows = rows_for<D>(sh, ws, axis, n);
                double* M = ws + sh.lay.base[p];
                s = transform_axis<D>(cx, M, sh.shape[p], sh.stride[p], axis, ws + sh.lay.cmat[axis], sh.mat_n[axis], rows,
                                      M, sh.stride[p], true, prm.use_fma != 0);
                if (cx.tid == 0) {
                    unsigned long long fibers = 1;
                    for (int a = 0; a < D; ++a) if (a != axis) fibers *= (unsigned long long)sh.shape[p][a];
                    sh.c_macs += fibers * ((unsigned long long)rows * n - (unsigned long long)rows * (rows - 1) / 2);
                    sh.shape[p][axis] = rows;
                }
                cx.sync();
            }
        }
        if (cx.tid == 0) sh.sumabs[p] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// Level-0 boxes.  A job is (system, optional restriction alpha/beta per axis, tracked interval,
// level, next split points): the whole unit cube for a fresh solve, a sub-box when the host
// re-solves merged exterior boxes (ChebyshevSubdivisionSolver.py:1376-1377) -- and, used on its
// own, the batched form of transformChebToInterval / TransformChebInPlaceND.
template <int D>
struct InitArgs {
    const double* coeffs; const long long* coeff_off;   // [nsys][D] offsets (doubles) of each tensor
    const int* shapes;                                   // [nsys][D][D]
    const double* errs;                                  // [nsys][D]
    const int* job_system;                               // [njobs]
    const int* job_restrict;                             // [njobs] 0: copy, 1: restrict by job_alpha/beta
    const double* job_alpha; const double* job_beta;     // [njobs][D]
    const double* job_iv;                                // [njobs][D][2] tracked interval of the new box
    const int* job_level;                                // [njobs]
    const double* job_next_mid;                          // [njobs][D]
    const int* job_parent;                               // [njobs] NodeRec index or -1
    unsigned long long njobs;
    BoxRec<D>* recs_out; double* data_out; unsigned long long cap_recs_out, cap_data_out;
    LevelCounters* ctr;
    double* gws; unsigned long long gws_stride; unsigned long long ws_doubles;
    SolverParams prm;
};

template <int D, class Ctx>
YR_DEV void init_worker(Ctx& cx, const InitArgs<D>& args, BoxShared<D>& sh, double* ws) {
    if (cx.tid == 0) { sh.c_macs = 0; }
    for (;;) {
        cx.sync();
        if (cx.tid == 0) sh.work_index = cx.atomic_add(&args.ctr->next_box, 1ull);
        cx.sync();
        const unsigned long long j = sh.work_index;
        if (j >= args.njobs) break;
        const int sys = args.job_system[j];
        if (cx.tid == 0) {
            for (int p = 0; p < D; ++p) {
                int st = 1;
                for (int a = D - 1; a >= 0; --a) { sh.shape[p][a] = args.shapes[(sys * D + p) * D + a]; sh.stride[p][a] = st; st *= sh.shape[p][a]; }
                sh.err[p] = args.errs[sys * D + p];
            }
            compute_layout<D>(sh, args.prm.exact != 0);
            IntervalState<D>& st = sh.rec.st;
            st.init_unit();
            for (int d = 0; d < D; ++d) {
                st.iv[d][0] = args.job_iv[(j * D + d) * 2]; st.iv[d][1] = args.job_iv[(j * D + d) * 2 + 1];
                st.next_mid[d] = args.job_next_mid[j * D + d];
                // local->top map of the new box: centre and half width of its interval, as double-double
                dd w = two_sum(st.iv[d][1], -st.iv[d][0]); dd c = two_sum(st.iv[d][1], st.iv[d][0]);
                st.A[d].hi = w.hi / 2; st.A[d].lo = w.lo / 2; st.B[d].hi = c.hi / 2; st.B[d].lo = c.lo / 2;
                sh.alpha[d] = args.job_alpha[j * D + d]; sh.beta[d] = args.job_beta[j * D + d];
                int n = 1;
                for (int p = 0; p < D; ++p) n = sh.shape[p][d] > n ? sh.shape[p][d] : n;
                sh.mat_axis[d] = d; sh.mat_n[d] = n;
            }
            sh.nmat = D;
        }
        cx.sync();
        for (int p = 0; p < D; ++p)
            cx.copy_in(ws + sh.lay.base[p], args.coeffs + args.coeff_off[sys * D + p], tensor_size<D>(sh.shape[p]));
        cx.sync();
        if (args.job_restrict[j]) {
            for (int p = 0; p < D; ++p) {
                const double s = tensor_abs_sum<D>(cx, ws + sh.lay.base[p], sh.shape[p], sh.stride[p]);
                if (cx.tid == 0) sh.sumabs[p] = s;
            }
            cx.sync();
            restrict_system_inplace<D>(cx, sh, ws, args.prm);
        }
        cx.sync();
        if (cx.tid == 0) {
            unsigned long long total = 0, widest = 0, fattest = 0;
            for (int p = 0; p < D; ++p) {
                const unsigned long long tsz = (unsigned long long)tensor_size<D>(sh.shape[p]);
                total += tsz; fattest = tsz > fattest ? tsz : fattest;
                for (int a = 0; a < D; ++a) widest = (unsigned long long)sh.shape[p][a] > widest ? sh.shape[p][a] : widest;
            }
            const unsigned long long rb = cx.atomic_add(&args.ctr->n_children, 1ull);
            const unsigned long long db = cx.atomic_add(&args.ctr->data_used, (total + 1ull) & ~1ull);
            cx.atomic_max(&args.ctr->max_child_doubles, total);
            cx.atomic_max(&args.ctr->max_child_extent, widest);
            cx.atomic_max(&args.ctr->max_child_tensor, fattest);
            sh.child_rec_base = rb; sh.child_off[0] = db;
            sh.action = kActNone;
            if (rb + 1 > args.cap_recs_out || db + total > args.cap_data_out) {
                cx.atomic_add(&args.ctr->overflow, 1ull); sh.action = kActDiscard;
            } else {
                BoxRec<D>& out = args.recs_out[rb];
                out.data_off = db;
                for (int p = 0; p < D; ++p) { for (int a = 0; a < D; ++a) out.shape[p][a] = sh.shape[p][a]; out.err[p] = sh.err[p]; }
                out.st = sh.rec.st;
                out.system = sys; out.parent_node = args.job_parent[j]; out.collector = -1;
                out.level = args.job_level[j]; out.path_hi = 0; out.path_lo = 0; out.path_bits = 0; out.pad0 = 0;
            }
        }
        cx.sync();
        if (sh.action != kActDiscard) {
            double* dst = args.data_out + sh.child_off[0];
            for (int p = 0; p < D; ++p) {
                const int cnt = tensor_size<D>(sh.shape[p]);
                const double* M = ws + sh.lay.base[p];
                for (int e = cx.tid; e < cnt; e += cx.nthreads) dst[e] = M[view_offset<D>(e, sh.shape[p], sh.stride[p])];
                dst += cnt;
            }
        }
    }
    if (cx.tid == 0) cx.atomic_add(&args.ctr->restrict_macs, sh.c_macs);
}

}  // namespace yr
