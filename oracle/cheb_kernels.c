/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the
 * product; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * Plain-C restatement of the reference's numba-JIT'd scalar kernels, compiled
 * with -ffp-contract=off so that, like numba's machine code (no
 * vfmadd, no contract flags), every multiply and add is a separately rounded
 * IEEE-754 double operation.
 *
 * Restated functions (reference: /root/reference/yroots/ChebyshevSubdivisionSolver.py):
 *   orc_restrict_axis0          <- TransformChebInPlace1D               (:45-120)
 *   orc_restrict_axis0_exact    <- TransformChebInPlace1DErrorFree      (:122-238)
 *   orc_restrict_axis0_split    <- TransformChebInPlace1DErrorFreeSplit (:240-337)
 *   orc_linear_check            <- linearCheck1                         (:608-626)
 *   two_sum/veltkamp/two_prod_s <- TwoSum/Split/TwoProdWithSplit        (:755-805)
 *
 * Layout: `src` is a C-contiguous [n][m] block, the restricted axis first and
 * all other axes flattened into m.  `dst` is [n][m], zero-filled here; only
 * the first `rows` (the return value) rows are meaningful.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double hi, lo; } pair_t;

static inline pair_t two_sum(double a, double b) {          /* TwoSum :756-761 */
    pair_t r; r.hi = a + b;
    double z = r.hi - a;
    r.lo = (a - (r.hi - z)) + (b - z);
    return r;
}
static inline pair_t veltkamp(double a) {                    /* Split :770-775 */
    pair_t r; double c = 134217729.0 * a;                    /* 2^27 + 1 */
    r.hi = c - (c - a);
    r.lo = a - r.hi;
    return r;
}
/* TwoProdWithSplit :800-805 (a already split into a1 + a2). */
static inline pair_t two_prod_s(double a, double b, double a1, double a2) {
    pair_t r; r.hi = a * b;
    pair_t bs = veltkamp(b);
    r.lo = a2 * bs.lo - (((r.hi - a1 * bs.hi) - a2 * bs.hi) - a1 * bs.lo);
    return r;
}

/* out[row][:] += w * src_row[:]  */
static inline void axpy_row(double *out, const double *x, double w, long m) {
    for (long j = 0; j < m; ++j) out[j] += x[j] * w;
}

/*
 * Column k of the restriction matrix C(alpha,beta) holds the Chebyshev
 * coefficients of T_k(alpha*x+beta).  Three live columns prev/cur/next are
 * rotated; each finished entry is immediately folded into dst.
 */
long orc_restrict_axis0(const double *src, long n, long m,
                        double alpha, double beta, double *dst)
{
    memset(dst, 0, sizeof(double) * (size_t)n * (size_t)m);
    if (n <= 0) return 0;
    memcpy(dst, src, sizeof(double) * (size_t)m);            /* column 0 = e0 */
    if (n == 1) return 1;
    double *buf = (double *)calloc((size_t)(3 * n + 3), sizeof(double));
    double *prev = buf, *cur = buf + (n + 1), *next = buf + 2 * (n + 1);
    prev[0] = 1.0;
    cur[0] = beta; cur[1] = alpha;                           /* column 1 */
    axpy_row(dst, src + m, beta, m);
    axpy_row(dst + m, src + m, alpha, m);
    long rows = 2;
    for (long k = 2; k < n; ++k) {
        const double *x = src + k * m;
        next[0] = -prev[0] + alpha * cur[1] + 2 * beta * cur[0];
        axpy_row(dst, x, next[0], m);
        if (rows > 2) {
            next[1] = -prev[1] + alpha * (2 * cur[0] + cur[2]) + 2 * beta * cur[1];
            axpy_row(dst + m, x, next[1], m);
        }
        for (long i = 2; i < rows - 1; ++i) {
            next[i] = -prev[i] + alpha * (cur[i - 1] + cur[i + 1]) + 2 * beta * cur[i];
            axpy_row(dst + i * m, x, next[i], m);
        }
        long i = rows - 1;
        next[i] = -prev[i] + (i == 1 ? 2.0 : 1.0) * alpha * cur[i - 1] + 2 * beta * cur[i];
        axpy_row(dst + i * m, x, next[i], m);
        double tail = alpha * cur[i];
        if (fabs(tail) > 1e-16) {                            /* :109 */
            next[rows] = tail;
            axpy_row(dst + rows * m, x, tail, m);
            rows += 1;
        }
        double *t = prev; prev = cur; cur = next; next = t;
    }
    free(buf);
    return rows;
}

long orc_restrict_axis0_split(const double *src, long n, long m, double sgn, double *dst);

long orc_restrict_axis0_exact(const double *src, long n, long m,
                              double alpha, double beta, double *dst)
{
    if (alpha == 0.5 && fabs(beta) == 0.5)                   /* :143-144 */
        return orc_restrict_axis0_split(src, n, m, beta > 0 ? 1.0 : -1.0, dst);
    memset(dst, 0, sizeof(double) * (size_t)n * (size_t)m);
    if (n <= 0) return 0;
    memcpy(dst, src, sizeof(double) * (size_t)m);
    if (n == 1) return 1;
    double *buf = (double *)calloc((size_t)(6 * (n + 1)), sizeof(double));
    double *p = buf, *c = p + (n + 1), *q = c + (n + 1);     /* values  */
    double *pe = q + (n + 1), *ce = pe + (n + 1), *qe = ce + (n + 1); /* errors */
    pair_t as = veltkamp(alpha), bs = veltkamp(beta);
    p[0] = 1.0; c[0] = beta; c[1] = alpha;
    axpy_row(dst, src + m, beta, m);
    axpy_row(dst + m, src + m, alpha, m);
    long rows = 2;
    for (long k = 2; k < n; ++k) {
        const double *x = src + k * m;
        {   /* row 0 :172-181 */
            pair_t v1 = two_prod_s(beta, 2 * c[0], bs.hi, bs.lo);
            pair_t v2 = two_prod_s(alpha, c[1], as.hi, as.lo);
            pair_t v3 = two_sum(v1.hi, v2.hi);
            pair_t v4 = two_sum(v3.hi, -p[0]);
            q[0] = v4.hi;
            qe[0] = -pe[0] + alpha * ce[1] + 2 * beta * ce[0] + v1.lo + v2.lo + v3.lo + v4.lo;
            axpy_row(dst, x, q[0] + qe[0], m);
        }
        if (rows > 2) {   /* row 1 :185-194 */
            pair_t v1 = two_sum(2 * c[0], c[2]);
            pair_t v2 = two_prod_s(beta, 2 * c[1], bs.hi, bs.lo);
            pair_t v3 = two_prod_s(alpha, v1.hi, as.hi, as.lo);
            pair_t v4 = two_sum(v2.hi, v3.hi);
            pair_t v5 = two_sum(v4.hi, -p[1]);
            q[1] = v5.hi;
            qe[1] = -pe[1] + alpha * (2 * ce[0] + ce[2] + v1.lo) + 2 * beta * ce[1]
                    + v2.lo + v3.lo + v4.lo + v5.lo;
            axpy_row(dst + m, x, q[1] + qe[1], m);
        }
        for (long i = 2; i < rows - 1; ++i) {   /* :197-206 */
            pair_t v1 = two_sum(c[i - 1], c[i + 1]);
            pair_t v2 = two_prod_s(beta, 2 * c[i], bs.hi, bs.lo);
            pair_t v3 = two_prod_s(alpha, v1.hi, as.hi, as.lo);
            pair_t v4 = two_sum(v2.hi, v3.hi);
            pair_t v5 = two_sum(v4.hi, -p[i]);
            q[i] = v5.hi;
            qe[i] = -pe[i] + alpha * (ce[i - 1] + ce[i + 1] + v1.lo) + 2 * beta * ce[i]
                    + v2.lo + v3.lo + v4.lo + v5.lo;
            axpy_row(dst + i * m, x, q[i] + qe[i], m);
        }
        {   /* second to last :209-218 */
            long i = rows - 1;
            double c1 = (i == 1) ? 2.0 : 1.0;
            pair_t v1 = two_prod_s(beta, 2 * c[i], bs.hi, bs.lo);
            pair_t v2 = two_prod_s(alpha, c1 * c[i - 1], as.hi, as.lo);
            pair_t v3 = two_sum(v1.hi, v2.hi);
            pair_t v4 = two_sum(v3.hi, -p[i]);
            q[i] = v4.hi;
            qe[i] = -pe[i] + c1 * alpha * ce[i - 1] + 2 * beta * ce[i]
                    + v1.lo + v2.lo + v3.lo + v4.lo;
            axpy_row(dst + i * m, x, q[i] + qe[i], m);
            /* last :222-227 -- always accumulated, grows only above 1e-32 */
            pair_t f = two_prod_s(alpha, c[i], as.hi, as.lo);
            qe[rows] = f.lo + alpha * ce[i];
            q[rows] = f.hi;
            axpy_row(dst + rows * m, x, q[rows] + qe[rows], m);
            if (fabs(q[rows] + qe[rows]) > 1e-32) rows += 1;
        }
        double *t = p; p = c; c = q; q = t;
        t = pe; pe = ce; ce = qe; qe = t;
    }
    free(buf);
    return rows;
}

long orc_restrict_axis0_split(const double *src, long n, long m, double sgn, double *dst)
{
    memset(dst, 0, sizeof(double) * (size_t)n * (size_t)m);
    if (n <= 0) return 0;
    memcpy(dst, src, sizeof(double) * (size_t)m);
    if (n == 1) return 1;
    double *buf = (double *)calloc((size_t)(6 * (n + 1)), sizeof(double));
    double *p = buf, *c = p + (n + 1), *q = c + (n + 1);
    double *pe = q + (n + 1), *ce = pe + (n + 1), *qe = ce + (n + 1);
    p[0] = 1.0; c[0] = sgn * 0.5; c[1] = 0.5;
    for (long j = 0; j < m; ++j) {                           /* :274-275 */
        dst[j] += sgn * src[m + j] / 2;
        dst[m + j] += src[m + j] / 2;
    }
    long rows = 2;
    for (long k = 2; k < n; ++k) {
        const double *x = src + k * m;
        {   /* :284-288 */
            pair_t v1 = two_sum(c[1] / 2, sgn * c[0]);
            pair_t v2 = two_sum(v1.hi, -p[0]);
            q[0] = v2.hi;
            qe[0] = -pe[0] + ce[1] / 2 + sgn * ce[0] + v1.lo + v2.lo;
            axpy_row(dst, x, q[0] + qe[0], m);
        }
        if (rows > 2) {   /* :291-298 */
            pair_t v1 = two_sum(c[0], c[2] / 2);
            pair_t v2 = two_sum(v1.hi, sgn * c[1]);
            pair_t v3 = two_sum(v2.hi, -p[1]);
            q[1] = v3.hi;
            qe[1] = -pe[1] + ce[0] + ce[2] / 2 + sgn * ce[1] + v1.lo + v2.lo + v3.lo;
            axpy_row(dst + m, x, q[1] + qe[1], m);
        }
        for (long i = 2; i < rows - 1; ++i) {   /* :301-308 */
            pair_t v1 = two_sum(c[i - 1], c[i + 1]);
            pair_t v2 = two_sum(v1.hi / 2, sgn * c[i]);
            pair_t v3 = two_sum(v2.hi, -p[i]);
            q[i] = v3.hi;
            qe[i] = -pe[i] + (ce[i - 1] + ce[i + 1] + v1.lo) / 2 + sgn * ce[i] + v2.lo + v3.lo;
            axpy_row(dst + i * m, x, q[i] + qe[i], m);
        }
        {   /* :311-326 */
            long i = rows - 1;
            double c1 = (i == 1) ? 1.0 : 0.5;
            pair_t v1 = two_sum(c1 * c[i - 1], sgn * c[i]);
            pair_t v2 = two_sum(v1.hi, -p[i]);
            q[i] = v2.hi;
            qe[i] = -pe[i] + c1 * ce[i - 1] + sgn * ce[i] + v1.lo + v2.lo;
            axpy_row(dst + i * m, x, q[i] + qe[i], m);
            q[rows] = c[i] / 2;
            qe[rows] = ce[i] / 2;
            axpy_row(dst + rows * m, x, q[rows] + qe[rows], m);
            if (fabs(q[rows] + qe[rows]) > 1e-32) rows += 1;
        }
        double *t = p; p = c; c = q; q = t;
        t = pe; pe = ce; ce = qe; qe = t;
    }
    free(buf);
    return rows;
}

/* linearCheck1 :608-626.  A is [d][d] row-major; lo/hi receive the bounds. */
void orc_linear_check(const double *total, const double *A, const double *consts,
                      long d, double *lo, double *hi)
{
    for (long j = 0; j < d; ++j) { lo[j] = -INFINITY; hi[j] = INFINITY; }
    for (long r = 0; r < d; ++r)
        for (long j = 0; j < d; ++j) {
            double a = A[r * d + j];
            if (a == 0.0) continue;
            double v1 = total[r] / fabs(a) - 1;
            double v2 = 2 * consts[r] / a;
            double l, h;
            if (v2 >= 0) { l = -v1; h = v1 - v2; }
            else         { l = -v2 - v1; h = v1; }
            if (l > lo[j]) lo[j] = l;
            if (h < hi[j]) hi[j] = h;
        }
}
