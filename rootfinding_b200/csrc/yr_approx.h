// Approximator kernels (ChebyshevApproximator.py:28-89, 288): separable building blocks of
// interval_approximate_nd, each acting along the middle axis of an [outer][n][inner] view so
// the host can sweep the axes of an n-D grid without transposes.
//
//   dct1_worker        values -> Chebyshev coefficients along one axis: SciPy's unnormalised
//                      DCT-I, divided by deg, first and last output plane halved (:78-82)
//   poly_eval_worker   Chebyshev (Clenshaw, polynomial.py:59-74) or power (Horner, :47-51)
//                      evaluation of a coefficient tensor along one axis at given abscissae
//   abs_mean_worker    per-index mean and max of |t| over the other axes (:72, :288)
//
// Written against the same Ctx interface as yr_box.h.
#pragma once
#include "yr_math.h"

#ifndef YR_DEV
#if defined(__CUDACC__)
#define YR_DEV __device__
#else
#define YR_DEV
#endif
#endif

namespace yr {

constexpr int kMaxApproxDim = 8;

struct Dct1Args { const double* values; double* coeffs; long long outer, npts, inner; };

// cta handles a contiguous range of (outer, inner) fibers; table[m] = cos(pi*m/N), m in [0, 2N)
// (fill = false: `table` is a global-memory table an earlier launch filled -- axes too long for shared memory)
template <class Ctx>
YR_DEV void dct1_worker(Ctx& cx, const Dct1Args& a, double* table, long long cta, long long nctas, bool fill = true) {
    const long long N = a.npts - 1;
    if (fill) {
        for (long long m = cx.tid; m < 2 * N; m += cx.nthreads) table[m] = cx.cospi_ratio(m, N);
        cx.sync();
    }
    const long long per_o = a.npts * a.inner;
    const long long total = a.outer * per_o;
    for (long long id = cta * cx.nthreads + cx.tid; id < total; id += nctas * cx.nthreads) {
        const long long o = id / per_o, r = id - o * per_o;
        const long long j = r / a.inner, i = r - j * a.inner;
        const double* v = a.values + o * per_o + i;
        // y_j = x_0 + (-1)^j x_N + 2 sum_{k=1}^{N-1} x_k cos(pi j k / N)
        double acc = 0.0;
        long long m = 0;                                    // (j*k) mod 2N, updated incrementally
        for (long long k = 1; k < N; ++k) {
            m += j; if (m >= 2 * N) m -= 2 * N;
            acc += v[k * a.inner] * table[m];
        }
        double y = v[0] + ((j & 1) ? -v[N * a.inner] : v[N * a.inner]) + 2.0 * acc;
        y /= (double)N;
        if (j == 0 || j == N) y /= 2;
        a.coeffs[id] = y;
    }
}

struct PolyEvalArgs { const double* coef; const double* x; double* out; long long outer, ncoef, npts, inner; int basis; };

template <class Ctx>
YR_DEV void poly_eval_worker(Ctx& cx, const PolyEvalArgs& a, long long cta, long long nctas) {
    const long long per_o = a.npts * a.inner;
    const long long total = a.outer * per_o;
    const long long n = a.ncoef;
    for (long long id = cta * cx.nthreads + cx.tid; id < total; id += nctas * cx.nthreads) {
        const long long o = id / per_o, r = id - o * per_o;
        const long long j = r / a.inner, i = r - j * a.inner;
        const double* c = a.coef + o * n * a.inner + i;
        const double x = a.x[j];
        double res;
        if (a.basis == 0) {                                 // chebval :59-74
            if (n == 1) res = c[0] + 0.0 * x;
            else if (n == 2) res = c[0] + c[a.inner] * x;
            else {
                const double x2 = 2 * x;
                double c0 = c[(n - 2) * a.inner], c1 = c[(n - 1) * a.inner];
                for (long long k = 3; k <= n; ++k) {
                    const double tmp = c0;
                    c0 = c[(n - k) * a.inner] - c1;
                    c1 = tmp + c1 * x2;
                }
                res = c0 + c1 * x;
            }
        } else {                                            // polyval :47-51
            double c0 = c[(n - 1) * a.inner];
            for (long long k = 2; k <= n; ++k) c0 = c[(n - k) * a.inner] + c0 * x;
            res = c0;
        }
        a.out[id] = res;
    }
}

// out[o][i][in] = sum_{k >= i} P[i][k] * c[o][k][in], accumulated from zero in ascending k with separately rounded
// multiply and add: the power -> Chebyshev basis change of MultiPower.to_cheb along one axis (polynomial.py:587-640,
// P[i][k] = coefficient of T_i in x^k), bit-identical to the reference's `cheb_coeffs[:k+1] += As * coeff[k]` loop.
struct UpperAxisArgs { const double* coef; const double* P; double* out; long long outer, n, inner; };

template <class Ctx>
YR_DEV void upper_axis_worker(Ctx& cx, const UpperAxisArgs& a, long long cta, long long nctas) {
    const long long per_o = a.n * a.inner;
    const long long total = a.outer * per_o;
    for (long long id = cta * cx.nthreads + cx.tid; id < total; id += nctas * cx.nthreads) {
        const long long o = id / per_o, r = id - o * per_o;
        const long long i = r / a.inner, in = r - i * a.inner;
        const double* c = a.coef + o * per_o + in;
        const double* p = a.P + i * a.n;
        double acc = 0.0;
        for (long long k = i; k < a.n; ++k) acc = acc + p[k] * c[k * a.inner];
        a.out[id] = acc;
    }
}

struct AbsMeanArgs { const double* t; double* mean; double* maxabs; long long outer, n, inner; };

// one CTA per index j of the middle axis
template <class Ctx>
YR_DEV void abs_mean_worker(Ctx& cx, const AbsMeanArgs& a, long long j) {
    const long long cnt = a.outer * a.inner;
    double s = 0.0, mx = 0.0;
    for (long long e = cx.tid; e < cnt; e += cx.nthreads) {
        const long long o = e / a.inner, i = e - o * a.inner;
        const double v = fabs(a.t[(o * a.n + j) * a.inner + i]);
        s += v; mx = v > mx ? v : mx;
    }
    s = cx.block_sum(s);
    mx = cx.block_max(mx);
    if (cx.tid == 0) { a.mean[j] = s / (double)cnt; if (a.maxabs) a.maxabs[j] = mx; }
}

// ---- coefficient-decay degree selection on the device (ChebyshevApproximator.py:91-172, 228-311) --------------------
// A probe of getChebyshevDegrees is a coefficient tensor `c` (row-major, extents shape[]) of which the reference keeps the
// sub-box keep[] (an axis the function is constant along is sampled with degree 1 and sliced back to one plane, :57,:85).
struct DecayTensor { const double* c; long long shape[kMaxApproxDim]; long long keep[kMaxApproxDim]; };

// block j: mean over the kept sub-box of |c| at index j of `axis` (np.average(|coeff|, other axes), :288)
template <class Ctx>
YR_DEV void subbox_axis_mean_worker(Ctx& cx, const DecayTensor& t, int dim, int axis, double* chunk, long long j) {
    long long cnt = 1;
    for (int d = 0; d < dim; ++d) if (d != axis) cnt *= t.keep[d];
    double s = 0.0;
    for (long long e = cx.tid; e < cnt; e += cx.nthreads) {
        long long rem = e, off = 0, stride = 1;
        for (int d = dim - 1; d >= 0; --d) {
            long long idx;
            if (d == axis) idx = j; else { idx = rem % t.keep[d]; rem /= t.keep[d]; }
            off += idx * stride; stride *= t.shape[d];
        }
        s += fabs(t.c[off]);
    }
    s = cx.block_sum(s);
    if (cx.tid == 0) chunk[j] = s / (double)cnt;
}

// grid-stride partial of max |c2 - c1| over the kept sub-box of c2, c1 taken as 0 outside its own kept sub-box
// (hasConverged :108-129); every CTA writes its maximum to part[cta]
template <class Ctx>
YR_DEV void maxdiff_worker(Ctx& cx, const DecayTensor& t2, const DecayTensor& t1, int dim, double* part, long long cta, long long nctas) {
    long long cnt = 1;
    for (int d = 0; d < dim; ++d) cnt *= t2.keep[d];
    double mx = 0.0;
    for (long long e = cta * cx.nthreads + cx.tid; e < cnt; e += nctas * cx.nthreads) {
        long long rem = e, off2 = 0, off1 = 0, s2 = 1, s1 = 1;
        bool inside = true;
        for (int d = dim - 1; d >= 0; --d) {
            const long long idx = rem % t2.keep[d]; rem /= t2.keep[d];
            off2 += idx * s2; s2 *= t2.shape[d];
            off1 += idx * s1; s1 *= t1.shape[d];
            inside = inside && idx < t1.keep[d];
        }
        const double v = fabs(t2.c[off2] - (inside ? t1.c[off1] : 0.0));
        mx = v > mx ? v : mx;               // (a NaN never compares greater: like np.max(...) < tol being False, see decide)
        if (v != v) mx = v;
    }
    mx = cx.block_max_nan(mx);
    if (cx.tid == 0) part[cta] = mx;
}

// grid-stride partial of max|v| (supNorm :72)
template <class Ctx>
YR_DEV void absmax_worker(Ctx& cx, const double* v, long long n, double* part, long long cta, long long nctas) {
    double mx = 0.0;
    for (long long e = cta * cx.nthreads + cx.tid; e < n; e += nctas * cx.nthreads) { const double a = fabs(v[e]); mx = a > mx ? a : mx; if (a != a) mx = a; }
    mx = cx.block_max_nan(mx);
    if (cx.tid == 0) part[cta] = mx;
}

// What the host reads back per probe: yr_decay_out (include/yroots_b200.h).
struct DecayOut { double supnorm, tol, started, converged, degree, eps, rho, reserved; };

YR_HD double max_nan(const double* p, long long n) { double m = 0.0; for (long long i = 0; i < n; ++i) { if (p[i] != p[i]) return p[i]; if (p[i] > m) m = p[i]; } return m; }

// thread-serial decisions.  mode 0: first probe of an iteration -- supNorm, tol, startedConverging (:91-106).
// mode 1: second probe -- supNorm2, tol from max(supNorm, supNorm2), hasConverged (:108-129), getFinalDegree (:131-172).
YR_HD void decay_decide(int mode, const double* chunk, long long n, const double* sup_part, long long nsup, const double* diff_part,
                        long long ndiff, double prev_supnorm, double rel_tol, double abs_tol, DecayOut* out) {
    const double sup = max_nan(sup_part, nsup);
    out->supnorm = sup;
    if (mode == 0) {
        const double tol = abs_tol + sup * rel_tol;
        out->tol = tol;
        bool all = true;
        for (long long i = (n > 5 ? n - 5 : 0); i < n; ++i) all = all && (chunk[i] < tol);
        out->started = all ? 1.0 : 0.0;
        out->converged = 0.0; out->degree = 0.0; out->eps = 0.0; out->rho = 0.0;
        return;
    }
    const double tol = abs_tol + (prev_supnorm > sup ? prev_supnorm : sup) * rel_tol;      // max(supNorm, supNorm2) :300
    out->tol = tol;
    out->started = 1.0;
    const double diff = max_nan(diff_part, ndiff);
    out->converged = (diff < tol) ? 1.0 : 0.0;
    // getFinalDegree :131-172
    const long long settled = (long long)(3 * (n - 1) / 4);
    double tail = 0.0;
    for (long long i = settled; i < n; ++i) if (chunk[i] > tail) tail = chunk[i];
    double eps = 2 * (kMachEps > tail ? kMachEps : tail);
    long long degree = 1, last = -1;
    for (long long i = 0; i < n; ++i) if (chunk[i] > eps) last = i;
    if (last >= 0) degree = last > 1 ? last : 1;
    bool tiny = true;
    for (long long i = 1; i < n; ++i) tiny = tiny && (chunk[i] < tol);
    if (tiny) degree = 0;
    long long peak = 0;
    for (long long i = 1; i < n; ++i) if (chunk[i] > chunk[peak]) peak = i;
    if (eps == 0) eps = chunk[peak] * 1e-24;
    out->degree = (double)degree; out->eps = eps;
    out->rho = pow(chunk[peak] / eps, 1.0 / (double)((degree - peak) + 1));
}

}  // namespace yr
