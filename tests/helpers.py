"""Shared helpers for the parity tests (CPU emulator tier and GPU tier use the same checks)."""
import os
import warnings

import numpy as np

from oracle import subdivision as osub
from oracle.gen_golden import rand_coeffs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT_TOL = 1e-10          # BASELINE.json north_star: roots within 1e-10 (max-norm) of the reference's output


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def unpack(z, prefix):
    return [z[f"{prefix}_{i}"] for i in range(int(z[prefix + "_n"]))]


def seeded_systems(dim, deg, N, kind="randn"):
    np.random.seed(2)                                   # /root/reference/tests/gen_random_tests.py:3
    batch = rand_coeffs(dim, deg, N, kind)
    return [[np.ascontiguousarray(m) for m in batch[s]] for s in range(N)]


def match_points(A, B, tol):
    """assert_same_points of the reference's tests/test_Combined_Solver.py:12-29 (max-norm)."""
    A = np.asarray(A, float).reshape(len(A), -1)
    B = np.asarray(B, float).reshape(len(B), -1)
    assert A.shape == B.shape, (A.shape, B.shape)
    used = np.zeros(len(B), dtype=bool)
    worst = 0.0
    for a in A:
        d = np.max(np.abs(B - a), axis=1)
        j = int(np.argmin(d))
        assert d[j] < tol, f"point {a} not matched, best distance {d[j]}"
        assert not used[j], "matched the same point twice"
        used[j] = True
        worst = max(worst, d[j])
    return worst


def oracle_solve(Ms, errs, **kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        trace = osub.Trace()
        roots = osub.solve_chebyshev_subdivision([m.copy() for m in Ms], np.array(errs, dtype=float), trace=trace, **kw)
    return np.asarray(roots, dtype=float).reshape(-1, len(Ms)), trace


def check_batch_against_oracle(solve_batch, systems, errs, engine=None, **kw):
    """Runs `solve_batch` and the oracle on the same systems; identical counts, roots within ROOT_TOL."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got, sess = solve_batch(systems, errs, engine=engine, return_session=True, **kw)
    boxes = 0
    worst = 0.0
    okw = {k: v for k, v in kw.items() if k in ("exact", "constant_check", "low_dim_quadratic_check", "all_dim_quadratic_check")}
    for Ms, e, g in zip(systems, errs, got):
        want, trace = oracle_solve(Ms, e, **okw)
        boxes += trace.boxes
        assert len(g) == len(want), (len(g), len(want))
        if len(want):
            worst = max(worst, match_points(want, g, ROOT_TOL))
    return sess, boxes, worst


def edge_case_systems():
    """(name, Ms, errors) inputs at the edges of the domain: constant / ragged / extent-1 tensors, one variable, no
    roots, huge error bounds, a NaN coefficient, an extent beyond the shared-memory pipeline (80 > 79)."""
    e1, e2 = np.full(1, 2.0 ** -52), np.full(2, 2.0 ** -52)
    rs = np.random.RandomState(5)
    return [
        ("constant 1x1", [np.array([[1.0]]), np.array([[2.0]])], e2),
        ("ragged 2x7 / 9x1", [np.random.RandomState(0).randn(2, 7), np.random.RandomState(1).randn(9, 1)], e2),
        ("linear", [np.array([[0.1, 0.0], [1.0, 0.0]]), np.array([[-0.2, 1.0], [0.0, 0.0]])], e2),
        ("no roots", [np.array([[5.0, 1.0], [1.0, 0.0]]), np.array([[0.0, 1.0], [1.0, 0.0]])], e2),
        ("1-D degree 30", [np.random.RandomState(3).randn(31)], e1),
        ("1-D constant", [np.array([3.0])], e1),
        ("extent 80", [rs.randn(80, 3), rs.randn(3, 80)], e2),
        ("huge error", [rs.randn(5, 5), rs.randn(5, 5)], np.full(2, 1e3)),
        ("nan coefficient", [np.array([[np.nan, 1.0], [1.0, 0.0]]), np.array([[0.0, 1.0], [1.0, 0.0]])], e2),
    ]


def check_edge_cases(S, osub):
    """Every edge case through solveChebyshevSubdivision `S` against the oracle: same root count, roots within ROOT_TOL."""
    import warnings
    for name, Ms, errs in edge_case_systems():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            got = np.asarray(S.solveChebyshevSubdivision([m.copy() for m in Ms], errs.copy())).reshape(-1, len(Ms))
            want = np.asarray(osub.solve_chebyshev_subdivision([m.copy() for m in Ms], errs.copy())).reshape(-1, len(Ms))
        assert got.shape == want.shape, (name, got.shape, want.shape)
        if len(want):
            match_points(want, got, ROOT_TOL)
    assert S.solve_batch([], []) == []
    assert S.solve_batch([], [], returnBoundingBoxes=True) == ([], [])


FIRST_SPLIT = 0.0394555475981047


def corner_system(seed, dim=2, deg=3):
    """A system with a root ON a corner of the subdivision grid: polynomial `ax` is (x_ax - m) * random, with m the
    reference's first split point (ChebyshevSubdivisionSolver.py:414).  All 2^dim level-1 cells report an exterior box at
    (m, ..., m); the reference merges them into ONE re-solve (:1323-1380)."""
    import numpy.polynomial.chebyshev as Ch
    rs = np.random.RandomState(seed)
    Ms = []
    for ax in range(dim):
        p = rs.randn(*([deg + 1] * dim))
        q = np.apply_along_axis(lambda c: Ch.chebmulx(c) - FIRST_SPLIT * np.r_[c, 0.0], ax, p)
        Ms.append(np.ascontiguousarray(q))
    return Ms


def match_boxes(want_roots, want_boxes, got_roots, got_boxes, tol, rel=0.05):
    """Pairs the reported bounding boxes by their centres and compares the endpoints (max-norm).  (Roots are not used for
    the pairing: a box with possibleDuplicateRoots reports several roots.)"""
    want_boxes, got_boxes = np.asarray(want_boxes, float), np.asarray(got_boxes, float)
    assert len(want_boxes) == len(got_boxes), (len(want_boxes), len(got_boxes))
    worst = 0.0
    used = np.zeros(len(got_boxes), dtype=bool)
    gc = got_boxes.mean(axis=2)
    for bx in want_boxes:
        d = np.max(np.abs(gc - bx.mean(axis=1)), axis=1)
        d[used] = np.inf
        j = int(np.argmin(d))
        used[j] = True
        diff = float(np.max(np.abs(got_boxes[j] - bx)))
        # a reported box is itself an error bound: endpoints agree to `tol` plus a few per cent of the reference box's width
        lim = tol + rel * float(np.max(bx[:, 1] - bx[:, 0]))
        assert diff < lim, f"bounding box {bx.tolist()} differs by {diff} (limit {lim})"
        worst = max(worst, diff)
    return worst


def sharded_cases():
    """(name, systems) inputs of the box-level multi-GPU tests (tests/test_distributed_gloo.py, test_gpu_parity)."""
    rs = np.random.RandomState(5)
    return [("one30", seeded_systems(2, 30, 1)),
            ("corners", seeded_systems(2, 12, 3) + [corner_system(s, 2, 3) for s in range(6)]),
            ("extent80", [[rs.randn(80, 3), rs.randn(3, 80)]])]


APPROX_FUNCS = {
    "cos": lambda x: np.cos(x),
    "exp_sin": lambda x, y: np.exp(x) * np.sin(3 * y),
    "readme_f": lambda x, y: np.sin(x * y) + x * np.log(y + 3) - x ** 2 + 1 / (y - 4),
    "readme_g": lambda x, y: np.cos(3 * x * y) + np.exp(3 * y / (x - 2)) - x - 6,
    "const_dim": lambda x, y, z: x ** 2 - y ** 2 + 3 * x * y,
    "poly3": lambda x, y, z: np.sin(5 * x + y + z),
}


def check_approx_error(err, z, name):
    """getApproxError (:313-356) against the reference's value.  The bound is built from epsVal = 2 max(2^-52, noise floor
    of the trailing axis means) and rho = (peak / epsVal)^(1/(deg - peak + 1)) (:131-172).  When the reference's epsVal sits
    on the 2^-52 floor on every axis the bound depends only on the peak means and must agree to 1e-9 (relative); when the
    rounding noise of the DCT itself sets epsVal (scipy's FFT vs the device's summation) only its size is comparable:
    within a factor 1.5."""
    want = float(z[f"{name}_err"])
    eps = np.asarray(z[f"{name}_eps"], dtype=float)
    on_floor = np.all((eps == 0) | (eps == 2.0 ** -51))
    if on_floor:
        assert abs(err - want) <= 1e-9 * abs(want), (name, err, want)
    else:
        assert 1 / 1.5 < (err + 1e-300) / (want + 1e-300) < 1.5, (name, err, want)
