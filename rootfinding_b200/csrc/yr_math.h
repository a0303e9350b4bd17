// Scalar per-box mathematics of the Chebyshev subdivision solver.
//
// Everything here is single-thread code usable from device kernels and from
// the host-side unit tests (tests/emul), which compile the same source with
// g++ -ffp-contract=off.  Device code is compiled with -fmad=false, so every
// a*b+c below is two separately rounded operations unless fma() is written
// explicitly -- the reference (numba, no FP contraction) rounds the
// same way.
//
// Reference behaviour followed (file:line in /root/reference/yroots/):
//   dd_* / two_sum / two_prod      ChebyshevSubdivisionSolver.py:755-805 (TwoSum/Split/TwoProd)
//   IntervalState, add_transform   ChebyshevSubdivisionSolver.py:376-576 (TrackedInterval)
//   linear_check                   ChebyshevSubdivisionSolver.py:608-626 (linearCheck1)
//   bounding_interval              ChebyshevSubdivisionSolver.py:628-753 (BoundingIntervalLinearSystem)
//   subdivision_axes               ChebyshevSubdivisionSolver.py:1013-1049 (getSubdivisionDims)
//   quad_check_2d/3d/nd            QuadraticCheck.py:34-687
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define YR_HD __host__ __device__ __forceinline__
#define YR_HDN static __host__ __device__ __noinline__   /* one copy per translation unit: instruction-cache footprint matters (profiles/) */
#else
#define YR_HD inline
#define YR_HDN inline
#endif
// Per-box mathematics with small array parameters (SVD, BILS, subdivision axes, 2-D quadratic check).  Out of line by
// default; a translation unit may define YR_HDM as force-inline so that the arrays live in registers instead of the
// stack (the 2-D frontier pipeline's thread-per-box scalar kernel does: its stack traffic was its bottleneck).
#ifndef YR_HDM
#define YR_HDM YR_HDN
#endif

namespace yr {

constexpr double kMachEps = 2.220446049250313e-16;       // 2^-52
constexpr double kFirstSplit = 0.0394555475981047;       // :414
constexpr int kMaxZoomCount = 25;                        // SolverOptions.maxZoomCount :39

// ---------------------------------------------------------------------------
// error-free transformations / double-double
// ---------------------------------------------------------------------------
struct dd { double hi, lo; };

// IEEE division / square root kept out of line on the device: every inline expansion is ~15-25
// instructions, and the scalar per-box code has dozens of them.
YR_HDN double yr_div(double a, double b) { return a / b; }
YR_HDN double yr_sqrt(double a) { return sqrt(a); }

YR_HD dd two_sum(double a, double b) {
    dd r; r.hi = a + b;
    double z = r.hi - a;
    r.lo = (a - (r.hi - z)) + (b - z);
    return r;
}
YR_HD dd quick_two_sum(double a, double b) {   // |a| >= |b|
    dd r; r.hi = a + b; r.lo = b - (r.hi - a); return r;
}
YR_HD dd two_prod(double a, double b) {        // exact product via one FMA
    dd r; r.hi = a * b; r.lo = fma(a, b, -r.hi); return r;
}
// Veltkamp split and Dekker product exactly as the reference writes them (exact mode only).
YR_HD dd veltkamp(double a) {
    dd r; double c = 134217729.0 * a; r.hi = c - (c - a); r.lo = a - r.hi; return r;
}
YR_HD dd two_prod_split(double a, double b, double a1, double a2) {
    dd r; r.hi = a * b;
    dd bs = veltkamp(b);
    r.lo = a2 * bs.lo - (((r.hi - a1 * bs.hi) - a2 * bs.hi) - a1 * bs.lo);
    return r;
}
YR_HD dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    dd t = two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = quick_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return quick_two_sum(s.hi, s.lo);
}
YR_HD dd dd_mul_d(dd a, double b) {
    dd p = two_prod(a.hi, b);
    p.lo += a.lo * b;
    return quick_two_sum(p.hi, p.lo);
}
YR_HD double dd_to_double(dd a) { return a.hi + a.lo; }

// ---------------------------------------------------------------------------
// per-box interval state (what TrackedInterval carries)
// ---------------------------------------------------------------------------
enum : uint32_t {
    kFlagFinalStep      = 1u << 0,   // TrackedInterval.finalStep
    kFlagCanThrowFinal  = 1u << 1,   // canThrowOutFinalStep
    kFlagEmpty          = 1u << 2,
};

template <int D>
struct IntervalState {
    double iv[D][2];        // current box in the solve's unit cube; plain doubles, exact at +-1 (:445-454)
    dd     A[D], B[D];      // local x in [-1,1] -> top coordinate A*x + B (composition of all (alpha,beta))
    double next_mid[D];     // nextTransformPoints
    double pf_iv[D][2];     // preFinalInterval
    dd     pfA[D], pfB[D];  // map at startFinalStep (-> reported bounding box)
    uint32_t flags;

    YR_HD bool final_step() const { return flags & kFlagFinalStep; }
    YR_HD bool can_throw_out() const {                     // :416-418
        return !(flags & kFlagFinalStep) || (flags & kFlagCanThrowFinal);
    }
    YR_HD bool is_point() const {                          // :558-560
        for (int d = 0; d < D; ++d) if (!(fabs(iv[d][0] - iv[d][1]) < 1e-32)) return false;
        return true;
    }
    YR_HD void start_final_step() {                        // :562-566
        flags |= kFlagFinalStep;
        for (int d = 0; d < D; ++d) {
            pf_iv[d][0] = iv[d][0]; pf_iv[d][1] = iv[d][1];
            pfA[d] = A[d]; pfB[d] = B[d];
        }
    }
    YR_HD void init_unit() {
        for (int d = 0; d < D; ++d) {
            iv[d][0] = -1.0; iv[d][1] = 1.0;
            A[d].hi = 1.0; A[d].lo = 0.0; B[d].hi = 0.0; B[d].lo = 0.0;
            next_mid[d] = kFirstSplit;
            pf_iv[d][0] = pf_iv[d][1] = 0.0;
            pfA[d] = A[d]; pfB[d] = B[d];
        }
        flags = 0;
    }
};

// addTransform :420-454.  `sub` is the sub-box in local [-1,1] coordinates.
// Returns false (and sets kFlagEmpty) when the box is inverted and may be
// thrown out.  alpha/beta receive the restriction applied to the tensors.
template <int D>
YR_HD bool add_transform(IntervalState<D>& s, double sub[D][2], double alpha[D], double beta[D]) {
    bool inverted = false;
    for (int d = 0; d < D; ++d) inverted = inverted || (sub[d][0] > sub[d][1]);
    if (inverted && s.can_throw_out()) { s.flags |= kFlagEmpty; return false; }
    if (inverted) {
        for (int d = 0; d < D; ++d) {
            double lo = fmax(fmin(sub[d][0], 1.0), -1.0);
            double hi = fmax(fmin(sub[d][1], 1.0), lo);
            sub[d][0] = lo; sub[d][1] = hi;
        }
    }
    for (int d = 0; d < D; ++d) {
        const double a1 = sub[d][0], b1 = sub[d][1];
        const double a2 = s.iv[d][0], b2 = s.iv[d][1];
        alpha[d] = (b1 - a1) / 2; beta[d] = (b1 + a1) / 2;
        const double alpha2 = (b2 - a2) / 2, beta2 = (b2 + a2) / 2;
        double nlo, nhi;
        if (a1 == -1.0) nlo = a2; else if (a1 == 1.0) nlo = b2; else nlo = alpha2 * a1 + beta2;
        if (b1 == -1.0) nhi = a2; else if (b1 == 1.0) nhi = b2; else nhi = alpha2 * b1 + beta2;
        s.iv[d][0] = nlo; s.iv[d][1] = nhi;
        // compose the local->top map: A*(alpha x + beta) + B
        dd nb = dd_add(dd_mul_d(s.A[d], beta[d]), s.B[d]);
        s.A[d] = dd_mul_d(s.A[d], alpha[d]);
        s.B[d] = nb;
    }
    return true;
}

// ---------------------------------------------------------------------------
// small dense linear algebra
// ---------------------------------------------------------------------------
// One-sided Jacobi SVD of a DxD matrix (replaces np.linalg.svd, :692).  On
// return sig[] holds the singular values (unsorted), W = U*diag(sig), V the
// right vectors.
template <int D>
YR_HDM void jacobi_svd(const double A[D][D], double W[D][D], double V[D][D], double sig[D]) {
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) { W[i][j] = A[i][j]; V[i][j] = (i == j) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < D - 1; ++p)
#pragma unroll
            for (int q = p + 1; q < D; ++q) {                       // (unrolled: every W / V index is a compile-time constant)
                double app = 0, aqq = 0, apq = 0;
#pragma unroll
                for (int i = 0; i < D; ++i) { app += W[i][p] * W[i][p]; aqq += W[i][q] * W[i][q]; apq += W[i][p] * W[i][q]; }
                if (apq == 0.0 || fabs(apq) <= 4e-16 * yr_sqrt(app * aqq)) continue;   // columns orthogonal to ~2 ulp
                rotated = true;
                double zeta = yr_div(aqq - app, 2.0 * apq);
                double t = yr_div(zeta >= 0 ? 1.0 : -1.0, fabs(zeta) + yr_sqrt(1.0 + zeta * zeta));
                double c = yr_div(1.0, yr_sqrt(1.0 + t * t)), s = c * t;
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    double wp = W[i][p], wq = W[i][q];
                    W[i][p] = c * wp - s * wq; W[i][q] = s * wp + c * wq;
                    double vp = V[i][p], vq = V[i][q];
                    V[i][p] = c * vp - s * vq; V[i][q] = s * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    for (int j = 0; j < D; ++j) {
        double s2 = 0;
        for (int i = 0; i < D; ++i) s2 += W[i][j] * W[i][j];
        sig[j] = yr_sqrt(s2);
    }
}

// linearCheck1 :608-626
template <int D>
YR_HD void linear_check(const double total[D], const double A[D][D], const double consts[D],
                        double lo[D], double hi[D]) {
    for (int j = 0; j < D; ++j) { lo[j] = -INFINITY; hi[j] = INFINITY; }
    for (int r = 0; r < D; ++r)
        for (int j = 0; j < D; ++j) {
            const double a = A[r][j];
            if (a == 0.0) continue;
            const double v1 = yr_div(total[r], fabs(a)) - 1;
            const double v2 = yr_div(2 * consts[r], a);
            double l, h;
            if (v2 >= 0) { l = -v1; h = v1 - v2; } else { l = -v2 - v1; h = v1; }
            lo[j] = fmax(lo[j], l);
            hi[j] = fmin(hi[j], h);
        }
}

YR_HD double pow2_floor_log2(double x) {   // 2^floor(log2(x)), x > 0 finite
    unsigned long long bits;
#if defined(__CUDA_ARCH__)
    bits = (unsigned long long)__double_as_longlong(x);
#else
    memcpy(&bits, &x, 8);
#endif
    if (bits & 0x7FF0000000000000ull) {                      // normal number: keep the exponent field only
        bits &= 0x7FF0000000000000ull;
#if defined(__CUDA_ARCH__)
        return __longlong_as_double((long long)bits);
#else
        double r; memcpy(&r, &bits, 8); return r;
#endif
    }
    return ldexp(1.0, ilogb(x));                             // subnormal
}

struct BilsOut { bool changed, should_stop, throw_out; };

// BoundingIntervalLinearSystem :628-753.
//   lin[p][d]  linear coefficient of polynomial p along axis d (getLinearTerms)
//   c0[p]      constant coefficient, sumabs[p] = sum|M_p|, errors[p] approximation error
template <int D>
YR_HDM BilsOut bounding_interval(const double lin[D][D], const double c0[D], const double sumabs[D],
                                 const double errors_in[D], bool final_step, double out[D][2]) {
    double A[D][D], consts[D], total[D], lsum[D], err[D], errors[D];
    const double shrink_floor = 0.99;
    double basecase_floor = 1.0;
    for (int d = 0; d < D; ++d) basecase_floor *= 0.4;       // 0.4**dim
    // python's 0.4**dim is pow(); repeated multiplication agrees to <= 1 ulp and is only a threshold
    for (int p = 0; p < D; ++p) {
        errors[p] = final_step ? 0.0 : errors_in[p];
        consts[p] = c0[p];
        total[p] = sumabs[p] + errors[p];
        double s = 0;
        for (int d = 0; d < D; ++d) { A[p][d] = lin[p][d]; s += fabs(lin[p][d]); }
        lsum[p] = s;
        err[p] = total[p] - fabs(consts[p]) - lsum[p];
    }
    for (int p = 0; p < D; ++p) {                              // row scaling :669-678
        double big = 0;
        for (int d = 0; d < D; ++d) big = fmax(big, fabs(A[p][d]));
        if (big > 0) {
            const double inv = yr_div(1.0, pow2_floor_log2(big));   // power of two: x * inv == x / s exactly
            for (int d = 0; d < D; ++d) A[p][d] *= inv;
            consts[p] *= inv; total[p] *= inv; lsum[p] *= inv; err[p] *= inv; errors[p] *= inv;
        }
    }
    double col_scale[D];
    for (int d = 0; d < D; ++d) {                              // column preconditioning :680-687
        col_scale[d] = 1.0;
        double big = 0;
        for (int p = 0; p < D; ++p) big = fmax(big, fabs(A[p][d]));
        if (big > 0) {
            const double s = yr_div(1.0, pow2_floor_log2(big));
            col_scale[d] = s;
            for (int p = 0; p < D; ++p) { total[p] += fabs(A[p][d]) * (s - 1); A[p][d] *= s; }
        }
    }
    double W[D][D], V[D][D], sig[D];
    jacobi_svd<D>(A, W, V, sig);
    double smax = 0, smin = INFINITY;
    for (int d = 0; d < D; ++d) { smax = fmax(smax, sig[d]); smin = fmin(smin, sig[d]); }
    BilsOut r;
    const bool force_stop_if_ill = final_step;
    if (!(smax > 0)) {
        // A == 0: the reference's condNum is 0/0 = NaN, every bound becomes NaN, nothing
        // compares true: changed = False, throwOut = False, should_stop = finalStep (:727,:753).
        for (int d = 0; d < D; ++d) { out[d][0] = NAN; out[d][1] = NAN; }
        r.changed = false; r.throw_out = false; r.should_stop = final_step;
        return r;
    }
    const double cond = yr_div(smin, smax);
    const bool well = cond > 1e-10;
    const double pad = fmax(cond, 2.0) * kMachEps;             // :696 (always 2^-51)
    double Ainv[D][D], center[D];
    if (well) {
        double inv_s2[D];
        for (int k = 0; k < D; ++k) inv_s2[k] = yr_div(1.0, sig[k] * sig[k]);
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j) {
                double s = 0;
                for (int k = 0; k < D; ++k) s += V[i][k] * inv_s2[k] * W[j][k];
                Ainv[i][j] = s;
            }
        for (int i = 0; i < D; ++i) {
            double s = 0;
            for (int j = 0; j < D; ++j) s += Ainv[i][j] * consts[j];
            center[i] = -s;
        }
    }
    double lo0[D], hi0[D];
    linear_check<D>(total, A, consts, lo0, hi0);
    const bool force_stop = force_stop_if_ill && !well;
    double keep_lo[D], keep_hi[D];
    for (int attempt = 0; attempt < 2; ++attempt) {
        double lo[D], hi[D];
        for (int d = 0; d < D; ++d) { lo[d] = lo0[d]; hi[d] = hi0[d]; }
        if (well) {
            for (int i = 0; i < D; ++i) {
                double w = 0;
                for (int j = 0; j < D; ++j) w += fabs(Ainv[i][j] * err[j]);
                lo[i] = fmax(center[i] - w, lo[i]);
                hi[i] = fmin(center[i] + w, hi[i]);
            }
        }
        for (int d = 0; d < D; ++d) { lo[d] = lo[d] * col_scale[d] - pad; hi[d] = hi[d] * col_scale[d] + pad; }
        bool thr = false;
        for (int d = 0; d < D; ++d) thr = thr || (lo[d] > hi[d]) || (lo[d] > 1) || (hi[d] < -1);
        for (int d = 0; d < D; ++d) {
            if (lo[d] < -1) lo[d] = -1;
            if (hi[d] < -1) hi[d] = -1;
            if (lo[d] > 1) lo[d] = 1;
            if (hi[d] > 1) hi[d] = 1;
        }
        if (!well)   // in-place aliasing of the reference (:702-725): pass 2 starts from pass 1's result
            for (int d = 0; d < D; ++d) { lo0[d] = lo[d]; hi0[d] = hi[d]; }
        double ratio = 1.0;
        for (int d = 0; d < D; ++d) ratio *= (hi[d] - lo[d]);
        ratio /= (double)(1 << D);
        bool changed;
        if (thr) changed = true;
        else if (attempt == 0) changed = ratio < shrink_floor;
        else changed = ratio < basecase_floor;
        if (attempt == 0 && changed) {
            for (int d = 0; d < D; ++d) { out[d][0] = lo[d]; out[d][1] = hi[d]; }
            r.changed = true; r.should_stop = force_stop; r.throw_out = thr;
            return r;
        }
        if (attempt == 0) {
            for (int d = 0; d < D; ++d) { keep_lo[d] = lo[d]; keep_hi[d] = hi[d]; err[d] = errors[d]; }
        } else {
            // ill conditioned: keep_* alias the arrays pass 2 just rewrote
            for (int d = 0; d < D; ++d) {
                out[d][0] = well ? keep_lo[d] : lo[d];
                out[d][1] = well ? keep_hi[d] : hi[d];
            }
            r.changed = false; r.throw_out = false;
            r.should_stop = changed ? force_stop : (well || force_stop);
            return r;
        }
    }
    r.changed = false; r.should_stop = false; r.throw_out = false;
    return r;   // unreachable
}

// getSubdivisionDims :1013-1049.  order[p][0..k-1] = axes to halve tensor p along, in
// order; returns k (identical for every polynomial).  shape[p][d] are the extents.
template <int D>
YR_HDM int subdivision_axes(const int shape[D][D], const double iv[D][2], int level, int order[D][D]) {
    // The candidate list of the reference is always in ascending axis order, so a bit mask carries it; every loop below
    // runs over compile-time axis indices (no array is indexed by a run-time value: the thread-per-box kernels keep
    // everything in registers instead of a stack frame).
    unsigned mask = (1u << D) - 1u;
    int nc = D;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        const double a = iv[i][0], b = iv[i][1];
        const bool close = fabs(a - b) <= (1e-8 + 1e-5 * fabs(b));     // np.isclose defaults
        if (close && nc != 1) { mask &= ~(1u << i); nc -= 1; }
    }
    if (level <= 5) {
        double widest = -INFINITY;
#pragma unroll
        for (int i = 0; i < D; ++i) if ((mask >> i) & 1u) widest = fmax(widest, iv[i][1] - iv[i][0]);
        const double bar = yr_div(widest, 5.0);
#pragma unroll
        for (int i = 0; i < D; ++i) if (((mask >> i) & 1u) && !((iv[i][1] - iv[i][0]) > bar)) { mask &= ~(1u << i); nc -= 1; }
        if (nc > 1) {
            int per_axis[D]; int total = 0;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                per_axis[d] = 0;
#pragma unroll
                for (int p = 0; p < D; ++p) per_axis[d] += shape[p][d];
                total += per_axis[d];
            }
            const int floor_share = total / (D + 1);
            const unsigned snapshot = mask;                            // :1046 iterates over a copy
#pragma unroll
            for (int i = 0; i < D; ++i)
                if (((snapshot >> i) & 1u) && nc > 1 && per_axis[i] < floor_share) { mask &= ~(1u << i); nc -= 1; }
        }
    }
    // per polynomial: candidates by descending extent.  np.argsort (insertion sort for tiny arrays: stable) ascending, then
    // reversed: ties end up in DESCENDING axis order.  Slot j takes the candidate that j others precede.
#pragma unroll
    for (int p = 0; p < D; ++p) {
#pragma unroll
        for (int j = 0; j < D; ++j) order[p][j] = 0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            if (!((mask >> i) & 1u)) continue;
            int rank = 0;
#pragma unroll
            for (int q = 0; q < D; ++q)
                if (q != i && ((mask >> q) & 1u) && (shape[p][q] > shape[p][i] || (shape[p][q] == shape[p][i] && q > i))) rank += 1;
#pragma unroll
            for (int j = 0; j < D; ++j) if (rank == j) order[p][j] = i;
        }
    }
    return nc;
}

// ---------------------------------------------------------------------------
// quadratic interval checks (QuadraticCheck.py)
// ---------------------------------------------------------------------------
struct RangeWitness {
    double bound; bool below, above;
    YR_HD bool feed(double v) { below = below || (v < bound); above = above || (v > -bound); return below && above; }
};

// c = {c00, c10, c01, c20, c11, c02};  sumabs = sum|M|;  returns true when no root is possible.
YR_HDM bool quad_check_2d(const double c[6], double sumabs, double tol) {
    double s = 0; for (int i = 0; i < 6; ++i) s = s + fabs(c[i]);
    RangeWitness r{sumabs - s + tol, false, false};
    const double k0 = c[0] - c[3] - c[5], k3 = 2 * c[3], k5 = 2 * c[5];
    auto q = [&](double x, double y) { return k0 + (c[1] + k3 * x + c[4] * y) * x + (c[2] + k5 * y) * y; };
    const double det = 16 * c[3] * c[5] - c[4] * c[4];
    double ix = INFINITY, iy = INFINITY;
    if (det != 0) { ix = yr_div(c[2] * c[4] - 4 * c[1] * c[5], det); iy = yr_div(c[1] * c[4] - 4 * c[2] * c[3], det); }
    if (r.feed(q(-1, -1))) return false;
    if (r.feed(q(1, -1))) return false;
    if (r.feed(q(-1, 1))) return false;
    if (r.feed(q(1, 1))) return false;
    if (c[5] != 0) {
        const double cc5 = 4 * c[5];
        for (int sx = -1; sx <= 1; sx += 2) {
            const double x = sx, y = yr_div(-(c[2] + c[4] * x), cc5);
            if (-1 < y && y < 1 && r.feed(q(x, y))) return false;
        }
    }
    if (c[3] != 0) {
        const double cc3 = 4 * c[3];
        for (int sy = -1; sy <= 1; sy += 2) {
            const double y = sy, x = yr_div(-(c[1] + c[4] * y), cc3);
            if (-1 < x && x < 1 && r.feed(q(x, y))) return false;
        }
    }
    if (-1 < ix && ix < 1 && -1 < iy && iy < 1 && r.feed(q(ix, iy))) return false;
    return true;
}

// c = {c000, c100, c010, c001, c110, c101, c011, c200, c020, c002}
YR_HDN bool quad_check_3d(const double c[10], double sumabs, double tol) {
    double s = 0; for (int i = 0; i < 10; ++i) s = s + fabs(c[i]);
    RangeWitness r{sumabs - s + tol, false, false};
    const double k0 = c[0] - c[7] - c[8] - c[9], k7 = 2 * c[7], k8 = 2 * c[8], k9 = 2 * c[9];
    auto q = [&](double x, double y, double z) {
        return k0 + (c[1] + k7 * x + c[4] * y + c[5] * z) * x + (c[2] + k8 * y + c[6] * z) * y + (c[3] + k9 * z) * z;
    };
    const double kk7 = 2 * k7, kk8 = 2 * k8, kk9 = 2 * k9;
    const double dx = kk8 * kk9 - c[6] * c[6], dy = kk7 * kk9 - c[5] * c[5], dz = kk7 * kk8 - c[4] * c[4];
    const double m12 = kk9 * c[4] - c[5] * c[6], m13 = c[4] * c[6] - kk8 * c[5], m23 = kk7 * c[6] - c[4] * c[5];
    const double det = 4 * c[7] * dx - c[4] * m12 + c[5] * m13;
    double ix = INFINITY, iy = INFINITY, iz = INFINITY;
    if (det != 0) {
        ix = yr_div((c[1] * -dx + c[2] * m12 + c[3] * -m13), det);
        iy = yr_div((c[1] * m12 + c[2] * -dy + c[3] * m23), det);
        iz = yr_div((c[1] * -m13 + c[2] * m23 + c[3] * -dz), det);
    }
    auto in = [](double v) { return -1 < v && v < 1; };
    const double L = -1, H = 1;
    if (r.feed(q(L, L, L))) return false;
    if (r.feed(q(H, L, L))) return false;
    if (r.feed(q(L, H, L))) return false;
    if (r.feed(q(L, L, H))) return false;
    if (r.feed(q(H, H, L))) return false;
    if (r.feed(q(H, L, H))) return false;
    if (r.feed(q(L, H, H))) return false;
    if (r.feed(q(H, H, H))) return false;
    if (c[9] != 0)
        for (int a = 0; a < 2; ++a) { const double x = a ? H : L, t = c[5] * x + c[3];
            for (int b = 0; b < 2; ++b) { const double y = b ? H : L, z = yr_div(-(t + c[6] * y), kk9);
                if (in(z) && r.feed(q(x, y, z))) return false; } }
    if (c[8] != 0)
        for (int a = 0; a < 2; ++a) { const double x = a ? H : L, t = c[2] + c[4] * x;
            for (int b = 0; b < 2; ++b) { const double z = b ? H : L, y = yr_div(-(t + c[6] * z), kk8);
                if (in(y) && r.feed(q(x, y, z))) return false; } }
    if (c[7] != 0)
        for (int a = 0; a < 2; ++a) { const double y = a ? H : L, t = c[1] + c[4] * y;
            for (int b = 0; b < 2; ++b) { const double z = b ? H : L, x = yr_div(-(t + c[5] * z), kk7);
                if (in(x) && r.feed(q(x, y, z))) return false; } }
    if (dx != 0)
        for (int a = 0; a < 2; ++a) { const double x = a ? H : L, u = c[2] + c[4] * x, v = c[3] + c[5] * x;
            const double y = yr_div((-kk9 * u + c[6] * v), dx), z = yr_div((c[6] * u - kk8 * v), dx);
            if (in(y) && in(z) && r.feed(q(x, y, z))) return false; }
    if (dy != 0)
        for (int a = 0; a < 2; ++a) { const double y = a ? H : L, u = c[1] + c[4] * y, v = c[3] + c[6] * y;
            const double x = yr_div((-kk9 * u + c[5] * v), dy), z = yr_div((c[5] * u - kk7 * v), dy);
            if (in(x) && in(z) && r.feed(q(x, y, z))) return false; }
    if (dz != 0)
        for (int a = 0; a < 2; ++a) { const double z = a ? H : L, u = c[1] + c[5] * z, v = c[2] + c[6] * z;
            const double x = yr_div((-kk8 * u + c[4] * v), dz), y = yr_div((c[4] * u - kk7 * v), dz);
            if (in(x) && in(y) && r.feed(q(x, y, z))) return false; }
    if (in(ix) && in(iy) && in(iz) && r.feed(q(ix, iy, iz))) return false;
    return true;
}

// Symmetric solve H x = rhs for an m x m system (m <= D) by Gaussian elimination with
// partial pivoting; returns false when numerically rank deficient (the reference asks
// np.linalg.matrix_rank(hermitian=True), :639/:672, and scipy.linalg.solve, :644/:673).
template <int D>
YR_HDN bool small_solve(int m, double H[D][D], double rhs[D], double x[D]) {
    double big = 0;
    for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) big = fmax(big, fabs(H[i][j]));
    if (!(big > 0)) return false;
    const double tiny = big * m * kMachEps;
    for (int k = 0; k < m; ++k) {
        int piv = k; double best = fabs(H[k][k]);
        for (int i = k + 1; i < m; ++i) if (fabs(H[i][k]) > best) { best = fabs(H[i][k]); piv = i; }
        if (!(best > tiny)) return false;
        if (piv != k) {
            for (int j = 0; j < m; ++j) { double t = H[k][j]; H[k][j] = H[piv][j]; H[piv][j] = t; }
            double t = rhs[k]; rhs[k] = rhs[piv]; rhs[piv] = t;
        }
        for (int i = k + 1; i < m; ++i) {
            const double f = yr_div(H[i][k], H[k][k]);
            for (int j = k; j < m; ++j) H[i][j] -= f * H[k][j];
            rhs[i] -= f * rhs[k];
        }
    }
    for (int i = m - 1; i >= 0; --i) {
        double s = rhs[i];
        for (int j = i + 1; j < m; ++j) s -= H[i][j] * x[j];
        x[i] = yr_div(s, H[i][i]);
    }
    return true;
}

// quadratic_check_nd :522-687 for any D (the default path for D == 1, and for D >= 4 on request).
//   c0 constant, g[d] linear, pure[d] coefficient of T_2(x_d), cross[i][j] (i<j) of x_i x_j;
//   rest = sum|M| - sum|those coefficients|.
template <int D>
YR_HDN bool quad_check_nd(double c0, const double g[D], const double pure[D], const double cross[D][D],
                          double rest, double tol) {
    double H[D][D], pure2[D];
    double psum = 0;
    for (int i = 0; i < D; ++i) { pure2[i] = pure[i] * 2; psum = psum + pure[i]; }
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j)
            H[i][j] = (i == j) ? pure2[i] * 2 : (i < j ? cross[i][j] : cross[j][i]);
    const double k0 = c0 - psum;
    auto q = [&](const double* pt) {
        double acc = k0;
        for (int i = 0; i < D; ++i) {
            double inner = 0;
            for (int j = i + 1; j < D; ++j) inner = inner + H[i][j] * pt[j];
            acc += (g[i] + pure2[i] * pt[i] + inner) * pt[i];
        }
        return acc;
    };
    RangeWitness r{rest + tol, false, false};
    double pt[D];
    for (int corner = 0; corner < (1 << D); ++corner) {          // itertools.product order: axis 0 most significant
        for (int i = 0; i < D; ++i) pt[i] = ((corner >> (D - 1 - i)) & 1) ? 1.0 : -1.0;
        if (r.feed(q(pt))) return false;
    }
    // subsets of fixed variables of size D-1 .. 1 (order of visiting does not change the answer)
    for (int mask = 1; mask < (1 << D) - 1; ++mask) {            // bit i set -> variable i is FREE
        int freev[D], fixv[D]; int nf = 0, nx = 0;
        for (int i = 0; i < D; ++i) { if ((mask >> i) & 1) freev[nf++] = i; else fixv[nx++] = i; }
        bool indefinite = false;
        for (int a = 0; a + 1 < nf; ++a) if (H[freev[a]][freev[a]] * H[freev[a + 1]][freev[a + 1]] < 0) indefinite = true;
        if (indefinite) continue;
        for (int side = 0; side < (1 << nx); ++side) {
            double Hs[D][D], rhs[D], xs[D];
            for (int a = 0; a < nf; ++a) {
                for (int b = 0; b < nf; ++b) Hs[a][b] = H[freev[a]][freev[b]];
                double s = 0;
                for (int b = 0; b < nx; ++b) s = s + H[freev[a]][fixv[b]] * (((side >> (nx - 1 - b)) & 1) ? 1.0 : -1.0);
                rhs[a] = -g[freev[a]] - s;
            }
            if (!small_solve<D>(nf, Hs, rhs, xs)) break;          // rank deficient for every side
            bool inside = true;
            for (int a = 0; a < nf; ++a) inside = inside && (-1.0 <= xs[a] && xs[a] <= 1.0);
            if (!inside) continue;
            for (int b = 0; b < nx; ++b) pt[fixv[b]] = ((side >> (nx - 1 - b)) & 1) ? 1.0 : -1.0;
            for (int a = 0; a < nf; ++a) pt[freev[a]] = xs[a];
            if (r.feed(q(pt))) return false;
        }
    }
    bool indefinite = false;
    for (int i = 0; i + 1 < D; ++i) if (pure[i] * pure[i + 1] < 0) indefinite = true;
    if (!indefinite) {
        double Hs[D][D], rhs[D], xs[D];
        for (int a = 0; a < D; ++a) { for (int b = 0; b < D; ++b) Hs[a][b] = H[a][b]; rhs[a] = -g[a]; }
        if (small_solve<D>(D, Hs, rhs, xs)) {
            bool inside = true;
            for (int a = 0; a < D; ++a) inside = inside && (-1.0 <= xs[a] && xs[a] <= 1.0);
            if (inside && r.feed(q(xs))) return false;
        }
    }
    return true;
}

}  // namespace yr
