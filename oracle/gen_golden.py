"""ORACLE -- test infrastructure only.

Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported in the build container with the one-line
`np.product = np.prod` shim numpy >= 2 needs) on fixed seeded inputs.
The reference cannot travel to the GPU box; these fixtures can.

    python oracle/gen_golden.py            # from the repo root, in the build container

Fixtures written (all float64 unless noted):
  restrict_axis.npz    inputs/outputs of TransformChebInPlaceND (plain + exact), incl. row counts
  bils.npz             inputs/outputs of BoundingIntervalLinearSystem (both passes, finalStep on/off)
  quadcheck.npz        inputs/outputs of quadratic_check (2-D, 3-D, n-D path)
  trim_subdivide.npz   inputs/outputs of trimMs, getSubdivisionIntervals, getInverseOrder, linearCheck1
  approx.npz           inputs/outputs of interval_approximate_nd / getFinalDegree / getApproxError / chebApproximate
  solves_random.npz    full-solve roots + box counts for seeded random systems (configs C3/C4)
  solves_chebfun.npz   yr.solve roots + box counts for the 27 Chebfun2 systems and the README system (C1/C2)
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
GOLD = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def import_reference():
    np.product = np.prod                                  # numpy>=2 shim
    sys.path.insert(0, "/root/reference")
    stub = types.ModuleType("matplotlib")
    stub.pyplot = types.ModuleType("matplotlib.pyplot")
    sys.modules.setdefault("matplotlib", stub)
    sys.modules.setdefault("matplotlib.pyplot", stub.pyplot)
    import yroots
    import yroots.ChebyshevApproximator as CA
    import yroots.ChebyshevSubdivisionSolver as CSS
    import yroots.QuadraticCheck as QC
    return yroots, CSS, CA, QC


def rand_coeffs(dim, deg, N, kind="randn", maxint=10):
    """Same stream as /root/reference/tests/random_tests.py:5-19 (call np.random.seed first)."""
    shape = (*(deg + 1,) * dim, dim, N)
    if kind == "randint":
        coeffs = np.random.randint(-maxint, maxint + 1, size=shape)
    else:
        coeffs = np.random.randn(*shape)
    for idx in np.ndindex((deg + 1,) * dim):
        if np.sum(idx) > deg:
            coeffs[idx] = 0
    return coeffs.T.astype(np.float64)


class BoxCounter:
    def __init__(self, CSS):
        self.CSS = CSS
        self.n = 0
        self._orig = CSS.solvePolyRecursive

        def counted(*a, **k):
            self.n += 1
            return self._orig(*a, **k)
        CSS.solvePolyRecursive = counted


def pack(prefix, arrays, out):
    out[prefix + "_n"] = np.array(len(arrays))
    for i, a in enumerate(arrays):
        out[f"{prefix}_{i}"] = np.asarray(a)


def main():
    os.makedirs(GOLD, exist_ok=True)
    os.chdir("/tmp")                                     # the reference may drop a debug file in cwd
    warnings.simplefilter("ignore")
    yroots, CSS, CA, QC = import_reference()
    counter = BoxCounter(CSS)
    rng = np.random.RandomState(20261019)

    # ---------------- restrict_axis -------------------------------------------------------
    out = {}
    cases = []
    shapes = [(7,), (33,), (6, 9), (21, 21), (51, 51), (4, 5, 6), (11, 11, 11), (3, 4, 5, 3), (5, 5, 5, 5)]
    abs_ = [(0.5, -0.5), (0.5, 0.5), (0.5197277737990523, -0.48027222620094767),
            (0.48027222620094767, 0.5197277737990523), (0.01, 0.3), (0.2, -0.7), (0.9, 0.05), (1e-6, 0.123)]
    k = 0
    for shp in shapes:
        M = rng.randn(*shp) * np.exp(-0.3 * np.indices(shp).sum(0))
        for axis in range(len(shp)):
            for (al, be) in abs_[(k % 3)::3] + abs_[:2]:
                for exact in (False, True):
                    R = CSS.TransformChebInPlaceND(M.copy(), axis, al, be, exact)
                    cases.append((M, axis, al, be, exact, np.ascontiguousarray(R)))
            k += 1
    out["n"] = np.array(len(cases))
    for i, (M, axis, al, be, exact, R) in enumerate(cases):
        out[f"in_{i}"] = M
        out[f"par_{i}"] = np.array([axis, al, be, float(exact)])
        out[f"out_{i}"] = R
    np.savez_compressed(os.path.join(GOLD, "restrict_axis.npz"), **out)
    print("restrict_axis:", len(cases), "cases")

    # ---------------- BILS ----------------------------------------------------------------
    out = {}
    n = 0
    for dim, deg in [(1, 12), (2, 4), (2, 9), (2, 20), (3, 5), (4, 3)]:
        for trial in range(10):
            decay = rng.choice([0.2, 0.8, 2.0])
            Ms = []
            for p in range(dim):
                shp = tuple(rng.randint(2, deg + 2, size=dim))
                M = rng.randn(*shp) * np.exp(-decay * np.indices(shp).sum(0))
                M.flat[0] *= rng.choice([0.01, 0.1, 1.0])
                if trial == 7:
                    M[(slice(1, 2),) + (0,) * (dim - 1)] = 0.0      # kill one linear term
                Ms.append(M)
            if trial == 8 and dim > 1:                               # singular linear part
                Ms[1] = Ms[0] * 1.5
            if trial == 9:                                           # all-zero linear part
                for M in Ms:
                    for d in range(dim):
                        idx = [0] * dim
                        idx[d] = 1
                        M[tuple(idx)] = 0.0
            errs = np.abs(rng.randn(dim)) * rng.choice([1e-15, 1e-8, 1e-3])
            for final in (False, True):
                iv, changed, stop, throw = CSS.BoundingIntervalLinearSystem([m.copy() for m in Ms], errs.copy(), final)
                pack(f"Ms_{n}", Ms, out)
                out[f"errs_{n}"] = errs
                out[f"final_{n}"] = np.array(final)
                out[f"iv_{n}"] = iv
                out[f"flags_{n}"] = np.array([changed, stop, throw], dtype=bool)
                n += 1
    out["n"] = np.array(n)
    np.savez_compressed(os.path.join(GOLD, "bils.npz"), **out)
    print("bils:", n, "cases")

    # ---------------- quadratic checks ----------------------------------------------------
    out = {}
    n = 0
    for dim, reps in [(1, 40), (2, 200), (3, 200), (4, 30)]:
        for trial in range(reps):
            shp = tuple(rng.randint(1, 6, size=dim))
            M = rng.randn(*shp) * np.exp(-rng.choice([0.5, 1.5, 3.0]) * np.indices(shp).sum(0))
            M.flat[0] = rng.randn() * rng.choice([0.1, 1.0, 3.0])
            tol = abs(rng.randn()) * rng.choice([1e-14, 1e-3])
            res = bool(QC.quadratic_check(M.copy(), tol))
            res_nd = bool(QC.quadratic_check(M.copy(), tol, nd_check=True))
            out[f"M_{n}"] = M
            out[f"tol_{n}"] = np.array(tol)
            out[f"res_{n}"] = np.array([res, res_nd])
            n += 1
    out["n"] = np.array(n)
    np.savez_compressed(os.path.join(GOLD, "quadcheck.npz"), **out)
    print("quadcheck:", n, "cases")

    # ---------------- trim / subdivide / helpers -------------------------------------------
    out = {}
    n = 0
    for dim, deg, level in [(1, 15, 2), (2, 12, 1), (2, 12, 7), (3, 6, 3), (3, 6, 8), (4, 3, 2)]:
        for trial in range(4):
            Ms = []
            for p in range(dim):
                shp = tuple(rng.randint(max(2, deg - 3), deg + 2, size=dim))
                M = rng.randn(*shp) * np.exp(-1.2 * np.indices(shp).sum(0))
                Ms.append(M)
            errs = np.abs(rng.randn(dim)) * 1e-6
            tM = [m.copy() for m in Ms]
            tE = errs.copy()
            CSS.trimMs(tM, tE)
            pack(f"Ms_{n}", Ms, out)
            out[f"errs_{n}"] = errs
            pack(f"trimM_{n}", [np.ascontiguousarray(m) for m in tM], out)
            out[f"trimE_{n}"] = tE
            box = CSS.TrackedInterval(np.array([[-1., 1.]] * dim))
            box = box.copy()
            if trial % 2 == 1:                       # a box that has been shrunk and split before
                sub = np.array([[-0.5, 0.25]] * dim)
                sub[0] = [-1.0, 0.3]
                box.addTransform(sub)
                box.nextTransformPoints[0] = 0
            iv_before = box.interval.copy()
            for exact in (False, True):
                cM, cE, cB = CSS.getSubdivisionIntervals([m.copy() for m in tM], tE.copy(), box.copy(), exact, level)
                tag = f"sub{int(exact)}_{n}"
                out[f"{tag}_nchild"] = np.array(len(cM))
                for c in range(len(cM)):
                    pack(f"{tag}_c{c}_M", [np.ascontiguousarray(m) for m in cM[c]], out)
                    out[f"{tag}_c{c}_E"] = np.array(cE[c])
                    out[f"{tag}_c{c}_iv"] = cB[c].interval
                    out[f"{tag}_c{c}_next"] = cB[c].nextTransformPoints
            out[f"box_iv_{n}"] = iv_before
            out[f"box_next_{n}"] = box.nextTransformPoints
            out[f"level_{n}"] = np.array(level)
            n += 1
    out["n"] = np.array(n)
    orders = [[0], [1, 0], [0, 1], [0, 3, 1], [2, 0, 1], [3, 1, 2, 0]]
    out["inv_n"] = np.array(len(orders))
    for i, o in enumerate(orders):
        out[f"inv_in_{i}"] = np.array(o)
        out[f"inv_out_{i}"] = np.array(CSS.getInverseOrder(np.array(o)))
    # linearCheck1 known-answer test of the reference's own old unit tests
    # (tests/_old_unit_tests/test_ChebyshevSubdivisionSolver.py:172-186) plus random ones
    lc = []
    for dim in (1, 2, 3, 5):
        for t in range(5):
            A = rng.randn(dim, dim)
            if t == 3:
                A[0, 0] = 0.0
            tot = np.abs(rng.randn(dim)) + 1.0
            c = rng.randn(dim) * 0.3
            lo, hi = CSS.linearCheck1(tot, A, c)
            lc.append((tot, A, c, lo, hi))
    out["lc_n"] = np.array(len(lc))
    for i, (tot, A, c, lo, hi) in enumerate(lc):
        out[f"lc_tot_{i}"], out[f"lc_A_{i}"], out[f"lc_c_{i}"] = tot, A, c
        out[f"lc_lo_{i}"], out[f"lc_hi_{i}"] = lo, hi
    np.savez_compressed(os.path.join(GOLD, "trim_subdivide.npz"), **out)
    print("trim/subdivide:", n, "cases")

    # ---------------- approximator ---------------------------------------------------------
    out = {}
    fns = {
        "cos": (lambda x: np.cos(x), [-1.], [1.]),
        "exp_sin": (lambda x, y: np.exp(x) * np.sin(3 * y), [-1., -2.], [0.5, 1.]),
        "readme_f": (lambda x, y: np.sin(x * y) + x * np.log(y + 3) - x ** 2 + 1 / (y - 4), [-1., -2.], [0., 1.]),
        "readme_g": (lambda x, y: np.cos(3 * x * y) + np.exp(3 * y / (x - 2)) - x - 6, [-1., -2.], [0., 1.]),
        "const_dim": (lambda x, y, z: x ** 2 - y ** 2 + 3 * x * y, [-1., -1., -1.], [1., 1., 1.]),
        "poly3": (lambda x, y, z: np.sin(5 * x + y + z), [-1., -1., -1.], [1., 1., 1.]),
    }
    names = sorted(fns)
    out["names"] = np.array(names)
    for name in names:
        f, a, b = fns[name]
        a, b = np.array(a), np.array(b)
        coeff, err = CA.chebApproximate(f, a, b)
        degs, eps, rhos = CA.getChebyshevDegrees(f, a, b, 1e-10)
        out[f"{name}_a"], out[f"{name}_b"] = a, b
        out[f"{name}_coeff"], out[f"{name}_err"] = coeff, np.array(err)
        out[f"{name}_degs"], out[f"{name}_eps"], out[f"{name}_rhos"] = degs.astype(float), eps, rhos
        probe = np.array([9] + [4] * (len(a) - 1))
        c2, sup = CA.interval_approximate_nd(f, probe.copy(), a, b, True)
        out[f"{name}_probe_degs"], out[f"{name}_probe_coeff"], out[f"{name}_probe_sup"] = probe, c2, np.array(sup)
    fd = []
    for t in range(12):
        m = rng.randint(10, 60)
        chunk = np.abs(rng.randn(m)) * np.exp(-rng.choice([0.3, 1.0, 3.0]) * np.arange(m)) + 1e-17 * rng.rand(m)
        tol = 1e-10 * chunk.max()
        d, e, r = CA.getFinalDegree(chunk, tol)
        fd.append((chunk, tol, d, e, r))
    out["fd_n"] = np.array(len(fd))
    for i, (chunk, tol, d, e, r) in enumerate(fd):
        out[f"fd_in_{i}"] = chunk
        out[f"fd_out_{i}"] = np.array([tol, d, e, r])
    np.savez_compressed(os.path.join(GOLD, "approx.npz"), **out)
    print("approx:", len(names), "functions")

    # ---------------- full solves on seeded random systems (C3 / C4) -----------------------
    out = {}
    plan = [(1, 10, 3, "randn"), (2, 3, 4, "randn"), (2, 5, 4, "randint"), (2, 10, 3, "randn"), (2, 20, 3, "randn"),
            (2, 30, 1, "randn"), (3, 3, 2, "randn"), (3, 6, 2, "randn"), (4, 2, 2, "randn"), (4, 3, 1, "randn"),
            (2, 50, 1, "randn")]
    n = 0
    for dim, deg, N, kind in plan:
        np.random.seed(2)                                       # gen_random_tests.py:3
        batch = rand_coeffs(dim, deg, N, kind)
        for s in range(N):
            Ms = [np.ascontiguousarray(m) for m in batch[s]]
            for exact in ((False, True) if deg <= 10 else (False,)):
                counter.n = 0
                roots = CSS.solveChebyshevSubdivision([m.copy() for m in Ms], np.full(dim, 2. ** -52), exact=exact)
                out[f"meta_{n}"] = np.array([dim, deg, s, float(kind == "randint"), float(exact), counter.n])
                out[f"roots_{n}"] = np.asarray(roots, dtype=float).reshape(-1, dim)
                n += 1
        print("  solved", dim, deg, N, kind)
    out["n"] = np.array(n)
    np.savez_compressed(os.path.join(GOLD, "solves_random.npz"), **out)
    print("solves_random:", n, "solves")

    # ---------------- Chebfun2 suite + README system through yr.solve (C1 / C2) ------------
    import chebfun_cases as cc
    out = {}
    for tid in cc.IDS:
        funcs, a, b = cc.system(tid)
        counter.n = 0
        roots, boxes = yroots.solve(funcs, a, b, returnBoundingBoxes=True)
        out[f"roots_{tid}"] = np.asarray(roots, dtype=float).reshape(-1, 2)
        out[f"boxes_{tid}"] = np.asarray(boxes, dtype=float).reshape(-1, 2, 2)
        out[f"nbox_{tid}"] = np.array(counter.n)
        if tid in ("2.3", "1.1", "6.3"):
            r2 = yroots.solve(funcs, a, b, exact=True)
            out[f"roots_exact_{tid}"] = np.asarray(r2, dtype=float).reshape(-1, 2)
    f = lambda x, y: np.sin(x * y) + x * np.log(y + 3) - x ** 2 + 1 / (y - 4)
    g = lambda x, y: np.cos(3 * x * y) + np.exp(3 * y / (x - 2)) - x - 6
    counter.n = 0
    roots = yroots.solve([f, g], [-1, -2], [0, 1])
    out["roots_readme"] = roots
    out["nbox_readme"] = np.array(counter.n)
    f1 = lambda x: np.sin(np.exp(3 * x))
    out["roots_univariate"] = np.asarray(yroots.solve(f1, -1, 2), dtype=float).reshape(-1, 1)
    np.savez_compressed(os.path.join(GOLD, "solves_chebfun.npz"), **out)
    print("solves_chebfun: done")


if __name__ == "__main__":
    main()
