// Thread-per-item bodies of the batched operator entry points (yr_bils_batched,
// yr_quadcheck_batched); shared by the CUDA backend and the test emulator.
#pragma once
#include "yr_box.h"

// ---- thread-per-item bodies shared by both backends ------------------------------------
namespace yr {

template <int D>
YR_HDN void bils_item(long long i, const double* lin, const double* c0, const double* sumabs, const double* errs,
                      const int* final_step, double* interval, int* flags) {
    double L[D][D], c[D], s[D], e[D], out[D][2];
    for (int p = 0; p < D; ++p) {
        c[p] = c0[i * D + p]; s[p] = sumabs[i * D + p]; e[p] = errs[i * D + p];
        for (int d = 0; d < D; ++d) L[p][d] = lin[(i * D + p) * D + d];
    }
    BilsOut r = bounding_interval<D>(L, c, s, e, final_step[i] != 0, out);
    for (int d = 0; d < D; ++d) { interval[(i * D + d) * 2] = out[d][0]; interval[(i * D + d) * 2 + 1] = out[d][1]; }
    flags[i * 3] = r.changed; flags[i * 3 + 1] = r.should_stop; flags[i * 3 + 2] = r.throw_out;
}

template <int D>
YR_HDN void quad_item(long long i, const double* coeffs, const long long* off, const int* shapes, const double* tol,
                      int nd_check, int* out) {
    // a one-tensor "box": reuse the gather + dispatch the solver kernel uses
    const double* M = coeffs + off[i];
    int shape[D], stride[D]; int st = 1;
    for (int a = D - 1; a >= 0; --a) { shape[a] = shapes[i * D + a]; stride[a] = st; st *= shape[a]; }
    double sumabs = 0;
    for (int e = 0; e < st; ++e) sumabs += fabs(M[e]);
    auto at = [&](int a, int va, int b, int vb) {
        int off2 = 0;
        for (int d = 0; d < D; ++d) {
            const int v = (d == a) ? va : ((d == b) ? vb : 0);
            if (v >= shape[d]) return 0.0;
            off2 += v * stride[d];
        }
        return M[off2];
    };
    bool res;
    if (D == 2 && !nd_check) {
        const double c[6] = {at(-1, 0, -1, 0), at(0, 1, -1, 0), at(1, 1, -1, 0), at(0, 2, -1, 0), at(0, 1, 1, 1), at(1, 2, -1, 0)};
        res = quad_check_2d(c, sumabs, tol[i]);
    } else if (D == 3 && !nd_check) {
        const double c[10] = {at(-1, 0, -1, 0), at(0, 1, -1, 0), at(1, 1, -1, 0), at(2, 1, -1, 0), at(0, 1, 1, 1),
                              at(0, 1, 2, 1), at(1, 1, 2, 1), at(0, 2, -1, 0), at(1, 2, -1, 0), at(2, 2, -1, 0)};
        res = quad_check_3d(c, sumabs, tol[i]);
    } else {
        double g[D], pure[D], cross[D][D];
        const double cz = at(-1, 0, -1, 0);
        double used = fabs(cz);
        for (int a = 0; a < D; ++a) { g[a] = at(a, 1, -1, 0); pure[a] = at(a, 2, -1, 0); used += fabs(g[a]) + fabs(pure[a]); }
        for (int a = 0; a < D; ++a) for (int b = 0; b < D; ++b) cross[a][b] = 0.0;
        for (int a = 0; a < D; ++a) for (int b = a + 1; b < D; ++b) { cross[a][b] = at(a, 1, b, 1); used += fabs(cross[a][b]); }
        res = quad_check_nd<D>(cz, g, pure, cross, sumabs - used, tol[i]);
    }
    out[i] = res ? 1 : 0;
}

}  // namespace yr
