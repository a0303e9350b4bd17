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

struct Dct1Args { const double* values; double* coeffs; long long outer, npts, inner; };

// cta handles a contiguous range of (outer, inner) fibers; table[m] = cos(pi*m/N), m in [0, 2N)
template <class Ctx>
YR_DEV void dct1_worker(Ctx& cx, const Dct1Args& a, double* table, long long cta, long long nctas) {
    const long long N = a.npts - 1;
    for (long long m = cx.tid; m < 2 * N; m += cx.nthreads) table[m] = cx.cospi_ratio(m, N);
    cx.sync();
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

}  // namespace yr
