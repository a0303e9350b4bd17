"""ORACLE -- test infrastructure only (see oracle/README.md).

CPU restatement of the reference's quadratic interval check
(/root/reference/yroots/QuadraticCheck.py:25-687).

A polynomial p = q + r is split into its total-degree<=2 Chebyshev part q and
the rest r.  If the range of q over [-1,1]^d stays outside
[-other_sum, other_sum], where other_sum = sum|r| + tol, then p has no zero on
the box.  The extrema of q are searched on corners, on edges/faces where some
partial derivatives vanish, and at the interior critical point.

The candidate points are generated with the same arithmetic expressions as the
reference so that the boolean outcome is reproducible; the answer itself does
not depend on the order in which candidates are visited.
"""
import itertools
from math import fabs, inf

import numpy as np
from scipy import linalg as sla


class _Range:
    """Tracks whether some value < other_sum and some value > -other_sum were seen."""
    __slots__ = ("bound", "below", "above")

    def __init__(self, bound):
        self.bound = bound
        self.below = False
        self.above = False

    def feed(self, v):
        """Returns True once both sides are witnessed (check is inconclusive)."""
        self.below = self.below or v < self.bound
        self.above = self.above or v > -self.bound
        return self.below and self.above


def quadratic_check(coeff, tol, nd_check=False):
    """Dispatcher, QuadraticCheck.py:25-31 (1-D and >=4-D go to the generic path)."""
    if coeff.ndim == 2 and not nd_check:
        return quadratic_check_2d(coeff, tol)
    if coeff.ndim == 3 and not nd_check:
        return quadratic_check_3d(coeff, tol)
    return quadratic_check_nd(coeff, tol)


def quadratic_check_2d(M, tol):
    """QuadraticCheck.py:34-182."""
    if M.ndim != 2:
        return False
    n0, n1 = M.shape
    c = [0] * 6
    c[0] = M[0, 0]
    if n0 > 1: c[1] = M[1, 0]
    if n1 > 1: c[2] = M[0, 1]
    if n0 > 2: c[3] = M[2, 0]
    if n0 > 1 and n1 > 1: c[4] = M[1, 1]
    if n1 > 2: c[5] = M[0, 2]
    other = np.sum(np.abs(M)) - sum([fabs(v) for v in c]) + tol      # :79
    k0 = c[0] - c[3] - c[5]
    k3 = 2 * c[3]
    k5 = 2 * c[5]

    def q(x, y):                                                      # :86-87
        return k0 + (c[1] + k3 * x + c[4] * y) * x + (c[2] + k5 * y) * y

    det = 16 * c[3] * c[5] - c[4] ** 2                                # :94
    if det != 0:
        ix = (c[2] * c[4] - 4 * c[1] * c[5]) / det
        iy = (c[1] * c[4] - 4 * c[2] * c[3]) / det
    else:
        ix = iy = inf
    r = _Range(other)
    for x, y in ((-1, -1), (1, -1), (-1, 1), (1, 1)):                 # corners :104-126
        if r.feed(q(x, y)):
            return False
    if c[5] != 0:                                                     # x fixed, dq/dy = 0
        cc5 = 4 * c[5]
        for x in (-1, 1):
            y = -(c[2] + c[4] * x) / cc5
            if -1 < y < 1 and r.feed(q(x, y)):
                return False
    if c[3] != 0:                                                     # y fixed, dq/dx = 0
        cc3 = 4 * c[3]
        for y in (-1, 1):
            x = -(c[1] + c[4] * y) / cc3
            if -1 < x < 1 and r.feed(q(x, y)):
                return False
    if -1 < ix < 1 and -1 < iy < 1 and r.feed(q(ix, iy)):             # interior :174
        return False
    return True


def quadratic_check_3d(M, tol):
    """QuadraticCheck.py:184-520."""
    if M.ndim != 3:
        return False
    s = M.shape
    c = [0] * 10
    c[0] = M[0, 0, 0]
    if s[0] > 1: c[1] = M[1, 0, 0]
    if s[1] > 1: c[2] = M[0, 1, 0]
    if s[2] > 1: c[3] = M[0, 0, 1]
    if s[0] > 1 and s[1] > 1: c[4] = M[1, 1, 0]
    if s[0] > 1 and s[2] > 1: c[5] = M[1, 0, 1]
    if s[1] > 1 and s[2] > 1: c[6] = M[0, 1, 1]
    if s[0] > 2: c[7] = M[2, 0, 0]
    if s[1] > 2: c[8] = M[0, 2, 0]
    if s[2] > 2: c[9] = M[0, 0, 2]
    other = np.sum(np.abs(M)) - sum([fabs(v) for v in c]) + tol      # :232
    k0 = c[0] - c[7] - c[8] - c[9]
    k7, k8, k9 = 2 * c[7], 2 * c[8], 2 * c[9]

    def q(x, y, z):                                                   # :240-243
        return k0 + (c[1] + k7 * x + c[4] * y + c[5] * z) * x + \
                    (c[2] + k8 * y + c[6] * z) * y + \
                    (c[3] + k9 * z) * z

    kk7, kk8, kk9 = 2 * k7, 2 * k8, 2 * k9
    dx = kk8 * kk9 - c[6] ** 2                                        # fix_x_det
    dy = kk7 * kk9 - c[5] ** 2
    dz = kk7 * kk8 - c[4] ** 2
    m12 = kk9 * c[4] - c[5] * c[6]
    m13 = c[4] * c[6] - kk8 * c[5]
    m23 = kk7 * c[6] - c[4] * c[5]
    det = 4 * c[7] * dx - c[4] * m12 + c[5] * m13
    if det != 0:
        ix = (c[1] * -dx + c[2] * m12 + c[3] * -m13) / det
        iy = (c[1] * m12 + c[2] * -dy + c[3] * m23) / det
        iz = (c[1] * -m13 + c[2] * m23 + c[3] * -dz) / det
    else:
        ix = iy = iz = inf
    lo, hi = -1, 1
    r = _Range(other)
    for p in ((lo, lo, lo), (hi, lo, lo), (lo, hi, lo), (lo, lo, hi),
              (hi, hi, lo), (hi, lo, hi), (lo, hi, hi), (hi, hi, hi)):  # :281-320
        if r.feed(q(*p)):
            return False
    inside = lambda v: lo < v < hi
    if c[9] != 0:                                                     # edges along z :324-356
        for x in (lo, hi):
            t = c[5] * x + c[3]
            for y in (lo, hi):
                z = -(t + c[6] * y) / kk9
                if inside(z) and r.feed(q(x, y, z)):
                    return False
    if c[8] != 0:                                                     # edges along y :361-393
        for x in (lo, hi):
            t = c[2] + c[4] * x
            for z in (lo, hi):
                y = -(t + c[6] * z) / kk8
                if inside(y) and r.feed(q(x, y, z)):
                    return False
    if c[7] != 0:                                                     # edges along x :398-430
        for y in (lo, hi):
            t = c[1] + c[4] * y
            for z in (lo, hi):
                x = -(t + c[5] * z) / kk7
                if inside(x) and r.feed(q(x, y, z)):
                    return False
    if dx != 0:                                                       # faces x = +-1 :436-456
        for x in (lo, hi):
            u = c[2] + c[4] * x
            v = c[3] + c[5] * x
            y = (-kk9 * u + c[6] * v) / dx
            z = (c[6] * u - kk8 * v) / dx
            if inside(y) and inside(z) and r.feed(q(x, y, z)):
                return False
    if dy != 0:                                                       # faces y = +-1 :462-482
        for y in (lo, hi):
            u = c[1] + c[4] * y
            v = c[3] + c[6] * y
            x = (-kk9 * u + c[5] * v) / dy
            z = (c[5] * u - kk7 * v) / dy
            if inside(x) and inside(z) and r.feed(q(x, y, z)):
                return False
    if dz != 0:                                                       # faces z = +-1 :488-508
        for z in (lo, hi):
            u = c[1] + c[5] * z
            v = c[2] + c[6] * z
            x = (-kk8 * u + c[4] * v) / dz
            y = (c[4] * u - kk7 * v) / dz
            if inside(x) and inside(y) and r.feed(q(x, y, z)):
                return False
    if inside(ix) and inside(iy) and inside(iz) and r.feed(q(ix, iy, iz)):
        return False
    return True


def quadratic_check_nd(M, tol):
    """QuadraticCheck.py:522-687 -- any dimension (used for 1-D and, on request, >=4-D)."""
    dim = M.ndim
    T = np.pad(M.copy(), [(0, max(0, 3 - n)) for n in M.shape], mode="constant")
    H = np.zeros((dim, dim))         # Hessian-like symmetric matrix of dq = H x + g
    g = np.zeros(dim)
    pure = [0] * dim
    const = 0.0
    for spot in itertools.product(range(3), repeat=dim):
        deg = sum(spot)
        if deg > 2:
            continue
        nz = [i for i in range(dim) if spot[i] != 0]
        if deg == 0:
            const = T[spot].copy()
        elif deg == 1:
            g[nz[0]] = T[spot].copy()
        elif len(nz) == 2:
            H[nz[1], nz[0]] = T[spot].copy()
            H[nz[0], nz[1]] = H[nz[1], nz[0]]
        else:
            pure[nz[0]] = T[spot].copy()
        T[spot] = 0
    pure2 = [p * 2 for p in pure]
    H[np.diag_indices(dim)] = [p * 2 for p in pure2]
    k0 = const - sum(pure)

    def q(pt):                                                        # :596-602
        acc = k0
        for i, x in enumerate(pt):
            acc += (g[i] + pure2[i] * x + sum([H[i, j] * pt[j] for j in range(i + 1, dim)])) * x
        return acc

    other = np.sum(np.abs(T)) + tol                                   # :605
    box = [-np.ones(dim), np.ones(dim)]
    r = _Range(other)
    for corner in itertools.product([0, 1], repeat=dim):              # :613-620
        if r.feed(q([box[j][i] for i, j in enumerate(corner)])):
            return False
    X = np.zeros(dim)
    fixed_sets = itertools.chain.from_iterable(
        itertools.combinations(range(dim), k) for k in range(dim - 1, 0, -1))   # :22-23
    for fixed in fixed_sets:
        fixed = np.array(fixed)
        free = np.delete(np.arange(dim), fixed)
        Hs = H[free][:, free]
        d = np.diag(Hs)
        if any(d[i] * d[i + 1] < 0 for i in range(len(d) - 1)):      # indefinite :632-635
            continue
        if np.linalg.matrix_rank(Hs, hermitian=True) != Hs.shape[0]:
            continue
        Hf = H[free][:, fixed]
        gs = g[free]
        for side in itertools.product([0, 1], repeat=len(fixed)):
            # NB (:643) the reference indexes the bound by the *position* in `fixed`;
            # every bound is +-1 so this is the same number.
            x0 = np.array([box[j][i] for i, j in enumerate(side)])
            xs = sla.solve(Hs, -gs - Hf @ x0, assume_a="sym")
            if all(box[0][v] <= xs[i] <= box[1][v] for i, v in enumerate(free)):
                X[fixed] = x0
                X[free] = xs
                if r.feed(q(X)):
                    return False
    if not any(pure[i] * pure[i + 1] < 0 for i in range(dim - 1)):    # interior :665-685
        if np.linalg.matrix_rank(H, hermitian=True) == H.shape[0]:
            X = sla.solve(H, -g, assume_a="sym")
            if all(box[0][i] <= X[i] <= box[1][i] for i in range(dim)):
                if r.feed(q(X)):
                    return False
    return True
