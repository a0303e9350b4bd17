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
