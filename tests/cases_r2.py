"""Round-2 parity inputs shared by oracle/gen_golden_r2.py (which runs the unmodified reference on them) and the tests.

Sources (all test DATA of the reference, restated as inputs):
  devastating(...)   /root/reference/tests/devastating_example_test_scripts.py:58-72, 93-110 (chebpolyqeps / chebspolyqeps)
  NOTEBOOK[...]      /root/reference/CombinedNotebook.ipynb cells 21 (Demo 1), 25 (Demo 2), 31 (Demo 4), 33 (3-D), 11 (5-D)
  FOUR_D[...]        /root/reference/tests/high_dim/4d_examples.py:95-465 (ex1, ex2, ex3, ex5, ex7, ex8, ex12)
"""
import numpy as np


def ortho_q(dim):
    from scipy.stats import ortho_group
    return ortho_group.rvs(dim, random_state=0)


def devastating_coeffs(kind, dim, eps):
    """Coefficient tensors (Chebyshev basis, shape [3]*dim) of the devastating example."""
    Q = ortho_q(dim)
    out = []
    if kind == "cheb":                                   # chebpolyqeps :58-72
        for i in range(dim):
            c = np.zeros([3] * dim)
            spot = [0] * dim
            c[tuple(spot)] = .5
            for j in range(dim):
                spot[j] = 1
                c[tuple(spot)] = Q[i, j] * eps
                spot[j] = 0
            spot[i] = 2
            c[tuple(spot)] = .5
            out.append(c)
    else:                                                # chebspolyqeps :93-110
        const = (1.5 * np.eye(dim) - eps * Q).sum(axis=1)
        for i in range(dim):
            c = np.zeros([3] * dim)
            spot = [0] * dim
            c[tuple(spot)] = const[i]
            for j in range(dim):
                spot[j] = 1
                c[tuple(spot)] = Q[i, j] * eps
                if i == j:
                    c[tuple(spot)] -= 2
                spot[j] = 0
            spot[i] = 2
            c[tuple(spot)] = .5
            out.append(c)
    return out


DEVASTATING = [(kind, dim, eps) for kind in ("cheb", "chebs") for dim in (2, 3) for eps in (1e-1, 1e-3, 1e-6)]


def dev_key(kind, dim, eps):
    return f"dev_{kind}_{dim}_{eps:g}"


NOTEBOOK = {
    "demo1": ([lambda x, y: np.sin(30 * x - y / 30) + y, lambda x, y: np.cos(x / 30 - 30 * y) - x], [-1, -1], [1, 1]),
    "demo2": ([lambda x, y: np.sin(20 * x + y), lambda x, y: np.cos(x ** 2 + x * y) - .25], [-5, -5], [5, 5]),
    "demo4": ([lambda x, y: 50 * np.cos(50 * x) * np.exp(np.sin(50 * x)) + 70 * np.cos(x) * np.cos(70 * np.sin(x)) - 10 * np.cos(10 * (x + y)) + .5 * x,
               lambda x, y: 60 * np.exp(y) * np.cos(60 * np.exp(y)) + 80 * np.cos(80 * y) * np.cos(np.sin(80 * y)) - 10 * np.cos(10 * (x + y)) + .5 * y],
              [-1, -1], [1, 1]),
    "demo3d": ([lambda x, y, z: np.sin(5 * x + y + z), lambda x, y, z: np.sin(x * y * z), lambda x, y, z: x ** 2 + y ** 2 - z ** 2 - 1],
               [-1, -1, -1], [1, 1, 1]),
    "demo5d": ([lambda x1, x2, x3, x4, x5: np.cos(x1) + x5 - 1, lambda x1, x2, x3, x4, x5: np.cos(x2) + x4 - 2,
                lambda x1, x2, x3, x4, x5: np.cos(x3) + x3 - 3, lambda x1, x2, x3, x4, x5: np.cos(x4) + x2 - 4,
                lambda x1, x2, x3, x4, x5: np.cos(x5) + x1 - 5], [0] * 5, [2 * np.pi] * 5),
}

FOUR_D = {
    "ex1": [lambda x1, x2, x3, x4: np.sin(x1 * x3) + x1 * np.log(x2 + 3) - x1 ** 2,
            lambda x1, x2, x3, x4: np.cos(4 * x1 * x2) + np.exp(3 * x2 / (x1 - 2)) - 5,
            lambda x1, x2, x3, x4: np.cos(2 * x2) - 3 * x3 + 1 / (x1 - 8),
            lambda x1, x2, x3, x4: x1 + x2 - x3 - x4],
    "ex2": [lambda x, y, z, x4: np.cosh(4 * x * y) + np.exp(z) - 5, lambda x, y, z, x4: x - np.log(1 / (y + 3)),
            lambda x, y, z, x4: x ** 2 - z, lambda x, y, z, x4: x + y + z + x4],
    "ex3": [lambda x, y, z, x4: y ** 2 - x ** 3, lambda x, y, z, x4: (y + .1) ** 3 - (x - .1) ** 2,
            lambda x, y, z, x4: x ** 2 + y ** 2 + z ** 2 - 1, lambda x, y, z, x4: x + y + z + x4],
    "ex5": [lambda x, y, z, x4: np.sin(4 * (x + z) * np.exp(y)), lambda x, y, z, x4: np.cos(2 * (z ** 3 + y + np.pi / 7)),
            lambda x, y, z, x4: 1 / (x + 5) - y, lambda x, y, z, x4: x - y + z - x4],
    "ex7": [lambda x, y, z, x4: np.cos(10 * x * y), lambda x, y, z, x4: x + y ** 2, lambda x, y, z, x4: x + y - z,
            lambda x, y, z, x4: x - y + z + x4],
    "ex8": [lambda x, y, z, x4: np.exp(2 * x) - 3, lambda x, y, z, x4: -np.exp(x - 2 * y) + 11, lambda x, y, z, x4: x + y + 3 * z,
            lambda x, y, z, x4: x + y + z + x4],
    "ex12": [lambda x, y, z, x4: x ** 2 + y ** 2 - .49 ** 2, lambda x, y, z, x4: (x - .1) * (x * y - .2),
             lambda x, y, z, x4: x ** 2 + y ** 2 - z ** 2, lambda x, y, z, x4: -x + y + z + x4],
}
FOUR_D_RESIDUAL = 2.22044604925031e-13                    # the threshold line of 4d_examples.py:42

# extra seeded random solves (same stream as tests/random_tests.py:5-19 after np.random.seed(2))
RANDOM_PLAN = [(2, 75, 1, "randn"), (2, 100, 1, "randn"), (3, 10, 3, "randn"), (3, 8, 2, "randint"),
               (4, 4, 2, "randn"), (4, 5, 1, "randn"), (5, 2, 2, "randn")]
# dimension-4 solves with all_dim_quadratic_check=True
ALLDIM_PLAN = [(4, 3, 2, "randn"), (4, 4, 1, "randn"), (5, 2, 1, "randn")]
