"""Synthetic workloads of the reference's own random tests, for benchmarks and profiling scripts.

`rand_coeffs` produces the coefficient tensors of /root/reference/tests/random_tests.py:5-19 (`rand_coeffs`, seeds and
degree tables in tests/gen_random_tests.py:3-27): i.i.d. N(0,1) (or integers in [-maxint, maxint]) coefficients on the
simplex sum(idx) <= deg, zero elsewhere, returned as ndarray[N, dim, deg+1, ..., deg+1] float64.  Call np.random.seed
first, as the reference's generator script does.
"""
import numpy as np


def rand_coeffs(dim, deg, N, kind="randn", maxint=10):
    shape = (*(deg + 1,) * dim, dim, N)
    if kind == "randint":
        coeffs = np.random.randint(-maxint, maxint + 1, size=shape).astype(np.float64)
    else:
        coeffs = np.random.randn(*shape)
    total = np.add.outer(np.arange(deg + 1), np.arange(deg + 1)) if dim == 2 else None
    if total is not None:
        coeffs[total > deg] = 0                                   # the simplex mask, all (dim, N) slices at once
    else:
        for idx in np.ndindex((deg + 1,) * dim):
            if np.sum(idx) > deg:
                coeffs[idx] = 0
    return coeffs.T.astype(np.float64)


def random_systems(dim, deg, N, seed=2, kind="randn"):
    """(systems, errors) for solve_batch: N systems of `dim` tensors, error bounds 2^-52 (Combined_Solver.py:122-124)."""
    np.random.seed(seed)
    arr = rand_coeffs(dim, deg, N, kind)
    systems = [[np.ascontiguousarray(m) for m in arr[s]] for s in range(N)]
    return systems, [np.full(dim, 2.0 ** -52)] * N


# ---- .npy test files of the reference's random-test harness (tests/random_tests.py:21-44) ----------------------------
FILEFMT = "tests/random_tests/dim{dim}_deg{deg}_{kind}.npy"      # random_tests.py:3


def gen_tests(dim, degrees, N, kind):
    """One coefficient array [N, dim, deg+1, ...] per degree (random_tests.py:21-25)."""
    return [rand_coeffs(dim, deg, N, kind) for deg in degrees]


def save_tests(dim, degrees, N, kind, filefmt=FILEFMT):
    """Writes the arrays of gen_tests in the reference's file layout (random_tests.py:27-31)."""
    import os
    for deg, arr in zip(degrees, gen_tests(dim, degrees, N, kind)):
        path = filefmt.format(dim=dim, deg=deg, kind=kind)
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        np.save(path, arr)


def load_tests(dim, deg, basis, kind, N=None, filefmt=FILEFMT):
    """Systems of Polynomial objects from a file written by save_tests or by the reference's own generator script
    (random_tests.py:32-44): a list of N lists of `dim` MultiPower (basis == "power") or MultiCheb objects."""
    from .polynomial import MultiCheb, MultiPower
    arr = np.load(filefmt.format(dim=dim, deg=deg, kind=kind))
    make = MultiPower if basis == "power" else MultiCheb
    if N is None:
        N = arr.shape[0]
    return [[make(coeff) for coeff in test] for test in arr[:N]]


def load_systems(dim, deg, kind, N=None, filefmt=FILEFMT):
    """The same file as (systems, errors) for solve_batch: raw float64 tensors, error bounds 2^-52."""
    arr = np.load(filefmt.format(dim=dim, deg=deg, kind=kind))
    if N is None:
        N = arr.shape[0]
    systems = [[np.ascontiguousarray(m, dtype=np.float64) for m in test] for test in arr[:N]]
    return systems, [np.full(dim, 2.0 ** -52)] * len(systems)
