"""Pins oracle/ (the CPU restatement) against outputs of the UNMODIFIED reference.

The fixtures under tests/golden/*.npz were produced by oracle/gen_golden.py,
which imports /root/reference in the build container; Polished_results/ and
Chebfun_results/ are the reference's own committed golden roots.  Everything
here runs on CPU (`-m "not gpu"`).
"""
import os
import warnings

import numpy as np
import pytest

import chebfun_cases as cc
import oracle
from oracle import approximator as oapp
from oracle import kernels as okern
from oracle import quadratic as oquad
from oracle import subdivision as osub
from oracle.gen_golden import rand_coeffs

warnings.simplefilter("ignore")


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _unpack(z, prefix):
    return [z[f"{prefix}_{i}"] for i in range(int(z[prefix + "_n"]))]


def test_restrict_axis_bit_exact(golden_dir):
    z = _load(golden_dir, "restrict_axis.npz")
    for i in range(int(z["n"])):
        axis, al, be, exact = z[f"par_{i}"]
        got = okern.restrict_axis(z[f"in_{i}"], int(axis), al, be, bool(exact))
        want = z[f"out_{i}"]
        assert got.shape == want.shape, (i, got.shape, want.shape)
        assert np.array_equal(np.ascontiguousarray(got), want), i


def test_bils_bit_exact(golden_dir):
    z = _load(golden_dir, "bils.npz")
    for i in range(int(z["n"])):
        Ms = _unpack(z, f"Ms_{i}")
        iv, changed, stop, throw = osub.bounding_interval_linear_system(
            [m.copy() for m in Ms], z[f"errs_{i}"].copy(), bool(z[f"final_{i}"]))
        assert np.array_equal(iv, z[f"iv_{i}"], equal_nan=True), i
        assert [bool(changed), bool(stop), bool(throw)] == list(z[f"flags_{i}"]), i


def test_quadratic_checks(golden_dir):
    z = _load(golden_dir, "quadcheck.npz")
    seen = {True: 0, False: 0}
    for i in range(int(z["n"])):
        M, tol = z[f"M_{i}"], float(z[f"tol_{i}"])
        res, res_nd = z[f"res_{i}"]
        assert bool(oquad.quadratic_check(M.copy(), tol)) == bool(res), i
        assert bool(oquad.quadratic_check(M.copy(), tol, nd_check=True)) == bool(res_nd), i
        seen[bool(res)] += 1
    assert seen[True] > 20 and seen[False] > 20      # both outcomes are exercised


def test_trim_and_subdivide_bit_exact(golden_dir):
    z = _load(golden_dir, "trim_subdivide.npz")
    for i in range(int(z["n"])):
        Ms = [m.copy() for m in _unpack(z, f"Ms_{i}")]
        errs = z[f"errs_{i}"].copy()
        osub.trim_system(Ms, errs)
        want = _unpack(z, f"trimM_{i}")
        assert all(np.array_equal(a, b) for a, b in zip(Ms, want)), i
        assert np.array_equal(errs, z[f"trimE_{i}"]), i
        dim = len(Ms)
        for exact in (0, 1):
            box = osub.Box(np.array([[-1., 1.]] * dim)).copy()
            box.interval = z[f"box_iv_{i}"].copy()
            box.nextTransformPoints = z[f"box_next_{i}"].copy()
            cM, cE, cB = osub.subdivide([m.copy() for m in Ms], errs.copy(), box, bool(exact), int(z[f"level_{i}"]))
            tag = f"sub{exact}_{i}"
            assert len(cM) == int(z[f"{tag}_nchild"])
            for c in range(len(cM)):
                for a, b in zip(cM[c], _unpack(z, f"{tag}_c{c}_M")):
                    assert np.array_equal(np.ascontiguousarray(a), b), (i, c)
                assert np.array_equal(np.array(cE[c]), z[f"{tag}_c{c}_E"]), (i, c)
                assert np.array_equal(cB[c].interval, z[f"{tag}_c{c}_iv"]), (i, c)
                assert np.array_equal(cB[c].nextTransformPoints, z[f"{tag}_c{c}_next"]), (i, c)
    for i in range(int(z["inv_n"])):
        assert tuple(z[f"inv_out_{i}"]) == osub.inverse_order(z[f"inv_in_{i}"])
    # the reference's docstring example, ChebyshevSubdivisionSolver.py:986-989
    assert osub.inverse_order(np.array([0, 3, 1])) == (0, 2, 1, 3, 4, 6, 5, 7)
    for i in range(int(z["lc_n"])):
        lo, hi = okern.linear_check(z[f"lc_tot_{i}"], z[f"lc_A_{i}"], z[f"lc_c_{i}"])
        assert np.array_equal(lo, z[f"lc_lo_{i}"]) and np.array_equal(hi, z[f"lc_hi_{i}"])


def test_linear_check_known_answer():
    """KAT from the reference's tests/_old_unit_tests/test_ChebyshevSubdivisionSolver.py:172-186."""
    A = np.array([[0.12122536, 0.58538202, 0.03219017],
                  [0.57397182, 0.51860304, 0.76381011],
                  [0.13666335, 0.29739732, 0.85181442]])
    totalErrs = np.array([1.61493601, 2.89650971, 2.18204467])
    consts = np.array([0.03935584, 0.08024904, 0.8969838])
    lo, hi = okern.linear_check(totalErrs, A, consts)
    # derived by hand from the definition: for each column the tightest row wins
    for col in range(3):
        cand_lo, cand_hi = [], []
        for row in range(3):
            v1 = totalErrs[row] / abs(A[row, col]) - 1
            v2 = 2 * consts[row] / A[row, col]
            cand_lo.append(-v1 if v2 >= 0 else -v2 - v1)
            cand_hi.append(v1 - v2 if v2 >= 0 else v1)
        assert lo[col] == max(cand_lo) and hi[col] == min(cand_hi)


def test_approximator(golden_dir):
    z = _load(golden_dir, "approx.npz")
    fns = {
        "cos": lambda x: np.cos(x),
        "exp_sin": lambda x, y: np.exp(x) * np.sin(3 * y),
        "readme_f": lambda x, y: np.sin(x * y) + x * np.log(y + 3) - x ** 2 + 1 / (y - 4),
        "readme_g": lambda x, y: np.cos(3 * x * y) + np.exp(3 * y / (x - 2)) - x - 6,
        "const_dim": lambda x, y, z: x ** 2 - y ** 2 + 3 * x * y,
        "poly3": lambda x, y, z: np.sin(5 * x + y + z),
    }
    for name in [str(s) for s in z["names"]]:
        f, a, b = fns[name], z[f"{name}_a"], z[f"{name}_b"]
        coeff, err = oapp.cheb_approximate(f, a, b)
        assert coeff.shape == z[f"{name}_coeff"].shape, name
        assert np.array_equal(coeff, z[f"{name}_coeff"]), name
        assert err == float(z[f"{name}_err"]), name
        degs, eps, rhos = oapp.chebyshev_degrees(f, a, b, 1e-10)
        assert np.array_equal(degs.astype(float), z[f"{name}_degs"])
        assert np.array_equal(eps, z[f"{name}_eps"]) and np.array_equal(rhos, z[f"{name}_rhos"])
        c2, sup = oapp.sample_and_dct(f, z[f"{name}_probe_degs"].copy(), a, b, True)
        assert np.array_equal(c2, z[f"{name}_probe_coeff"]) and sup == float(z[f"{name}_probe_sup"])
    for i in range(int(z["fd_n"])):
        tol, d, e, r = z[f"fd_out_{i}"]
        got = oapp.final_degree(z[f"fd_in_{i}"], tol)
        assert (float(got[0]), float(got[1]), float(got[2])) == (d, e, r)


def test_has_converged_truth_table():
    """The four cases of the reference's tests/_old_unit_tests/test_hasConverged.py:6-29."""
    tol = 0.01
    c = np.array([0.1, 1e-5, 0.3])
    assert not oapp.has_converged(c, np.array([0.1, 1e-5, 0.3, 0.5]), tol)          # new large tail entry
    assert oapp.has_converged(c, np.array([0.1, 1e-5, 0.3, 0.001]), tol)
    assert not oapp.has_converged(c, np.array([0.5, 1e-5, 0.3, 0.0]), tol)          # leading entry moved
    assert oapp.has_converged(c, np.array([0.1 + 1e-3, 1e-5, 0.3 - 1e-3, 0.0]), tol)


def test_full_solves_random_bit_exact(golden_dir):
    z = _load(golden_dir, "solves_random.npz")
    cache = {}
    for i in range(int(z["n"])):
        dim, deg, s, randint, exact, nbox = z[f"meta_{i}"]
        dim, deg, s = int(dim), int(deg), int(s)
        if deg >= 50:
            continue                                     # covered by the slower test below
        key = (dim, deg, bool(randint))
        if key not in cache:
            np.random.seed(2)
            N = {(1, 10): 3, (2, 3): 4, (2, 5): 4, (2, 10): 3, (2, 20): 3, (2, 30): 1,
                 (3, 3): 2, (3, 6): 2, (4, 2): 2, (4, 3): 1}[(dim, deg)]
            cache[key] = rand_coeffs(dim, deg, N, "randint" if randint else "randn")
        Ms = [np.ascontiguousarray(m) for m in cache[key][s]]
        trace = osub.Trace()
        roots = osub.solve_chebyshev_subdivision(Ms, np.full(dim, 2. ** -52), exact=bool(exact), trace=trace)
        want = z[f"roots_{i}"]
        assert np.asarray(roots).reshape(-1, dim).shape == want.shape, i
        assert np.array_equal(np.asarray(roots).reshape(-1, dim), want), i
        assert trace.boxes == int(nbox), (i, trace.boxes, nbox)


def test_full_solve_deg50_headline(golden_dir):
    z = _load(golden_dir, "solves_random.npz")
    idx = [i for i in range(int(z["n"])) if int(z[f"meta_{i}"][1]) == 50][0]
    np.random.seed(2)
    Ms = [np.ascontiguousarray(m) for m in rand_coeffs(2, 50, 1)[0]]
    trace = osub.Trace()
    roots = osub.solve_chebyshev_subdivision(Ms, np.full(2, 2. ** -52), trace=trace)
    assert np.array_equal(roots, z[f"roots_{idx}"])
    assert trace.boxes == int(z[f"meta_{idx}"][5])


@pytest.mark.parametrize("tid", cc.IDS)
def test_chebfun_suite_oracle(tid, golden_dir):
    """oracle.solve == reference solve bit for bit, and passes the suite's own criteria."""
    z = _load(golden_dir, "solves_chebfun.npz")
    funcs, a, b = cc.system(tid)
    trace = osub.Trace()
    roots, boxes = oracle.solve(funcs, a, b, returnBoundingBoxes=True, trace=trace)
    roots = np.asarray(roots, dtype=float).reshape(-1, 2)
    assert np.array_equal(roots, z[f"roots_{tid}"])
    assert np.array_equal(np.asarray(boxes, dtype=float).reshape(-1, 2, 2), z[f"boxes_{tid}"])
    assert trace.boxes == int(z[f"nbox_{tid}"])
    truth = cc.polished_roots(tid, golden_dir)
    if tid not in cc.COUNT_EXEMPT:
        assert len(roots) == len(truth) == cc.EXPECTED_COUNT[tid]
        ok, n0, n1 = cc.norm_check(roots, truth, cc.norm_tol(tid))
        assert ok, (tid, n0, n1)
    assert cc.residual_check(funcs, roots, cc.resid_tol(tid))
    if f"roots_exact_{tid}" in z.files:
        r2 = np.asarray(oracle.solve(funcs, a, b, exact=True), dtype=float).reshape(-1, 2)
        assert np.array_equal(r2, z[f"roots_exact_{tid}"])


def test_readme_and_univariate(golden_dir):
    z = _load(golden_dir, "solves_chebfun.npz")
    f = lambda x, y: np.sin(x * y) + x * np.log(y + 3) - x ** 2 + 1 / (y - 4)
    g = lambda x, y: np.cos(3 * x * y) + np.exp(3 * y / (x - 2)) - x - 6
    roots = oracle.solve([f, g], [-1, -2], [0, 1])
    assert np.array_equal(roots, z["roots_readme"]) and len(roots) == 2
    assert np.max(np.abs(f(roots[:, 0], roots[:, 1]))) < 1e-14        # tests/test_Combined_Solver.py:57-68
    assert np.max(np.abs(g(roots[:, 0], roots[:, 1]))) < 1e-14
    f1 = lambda x: np.sin(np.exp(3 * x))
    r1 = oracle.solve(f1, -1, 2)
    assert len(r1) == 128 and np.max(np.abs(f1(r1))) < 1e-12          # :32-36
    assert np.array_equal(np.asarray(r1).reshape(-1, 1), z["roots_univariate"])


def test_split_matrices_are_exact_mirror_images():
    """The two halves of a subdivision at m = 0 use C(1/2, -1/2) and C(1/2, +1/2); the frontier pipeline's
    mirrored sweep (csrc/yr_f2.h: pass_fiber_dual) relies on C_hi[i][k] == (-1)^(i+k) C_lo[i][k] holding bit for bit,
    rows included.  Checked here on the oracle's restatement of TransformChebInPlace1D (:45-120)."""
    from oracle import kernels as K
    n = 101
    eye = np.eye(n)
    lo = K.restrict_axis(eye, 0, 0.5, -0.5)
    hi = K.restrict_axis(eye, 0, 0.5, 0.5)
    assert lo.shape == hi.shape
    sign = (-1.0) ** (np.add.outer(np.arange(lo.shape[0]), np.arange(n)))
    assert np.array_equal(hi, lo * sign)
    # and a product shared between the halves rounds identically: x + (-(c*v)) == x - c*v
    rng = np.random.RandomState(1)
    c, v, x = rng.randn(1000), rng.randn(1000), rng.randn(1000)
    assert np.array_equal(x + (-c) * v, x - c * v)


def test_trim_policies_match_reference(golden_dir):
    """trimMs and trimMsOptimized1/2/3 of the reference's experiment file (tests/TrimMs optimization/...:1125-1271), run by
    oracle/gen_golden_trim.py: the oracle's restatements leave the same extents and the same errors, bit for bit."""
    z = _load(golden_dir, "trim_policies.npz")
    trimmed = 0
    for name in z["cases"]:
        dim = int(str(name)[1])
        Ms = [z[f"M_{name}_{p}"] for p in range(dim)]
        for k, fn in osub.TRIM_POLICIES.items():
            Mk, ek = [M.copy() for M in Ms], z[f"err_{name}"].copy()
            fn(Mk, ek)
            assert np.array_equal(np.array([M.shape for M in Mk]), z[f"shape_{name}_{k}"]), (name, k)
            assert np.array_equal(ek, z[f"err_{name}_{k}"]), (name, k)
            for p in range(dim):
                assert np.array_equal(Mk[p], Ms[p][tuple(slice(0, e) for e in Mk[p].shape)])
            trimmed += int(sum(M.size for M in Ms) - sum(M.size for M in Mk) > 0)
    assert trimmed >= 20
