"""ORACLE -- test infrastructure only (see oracle/README.md).

CPU restatement of the parts of /root/reference/yroots/polynomial.py that the
hot path touches: point-list evaluation of Chebyshev / power-basis coefficient
tensors (MultiCheb.__call__ :317-342, MultiPower.__call__ :503-528, helpers
chebval/polyval :47-91) and the power -> Chebyshev basis change
(MultiPower.to_cheb :587-640).
"""
import numpy as np


def _as_points(points, dim):
    """Polynomial.__call__ :176-205 -- normalise to an [npts, dim] array."""
    pts = np.array(points)
    if pts.ndim == 0:
        pts = np.array([pts])
    if pts.ndim == 1:
        pts = pts.reshape(1, pts.shape[0]) if dim > 1 else pts.reshape(pts.shape[0], 1)
    if pts.shape[1] != dim:
        raise ValueError('Dimension of points does not match dimension of polynomial!')
    return pts


def _clenshaw(x, cc):
    """chebval :59-74 -- Clenshaw recurrence along the leading axis of cc."""
    n = len(cc)
    if n == 1:
        c0, c1 = cc[0], np.zeros_like(cc[0])
    elif n == 2:
        c0, c1 = cc[0], cc[1]
    else:
        x2 = 2 * x
        c0, c1 = cc[-2], cc[-1]
        for i in range(3, n + 1):
            c0, c1 = cc[-i] - c1, c0 + c1 * x2
    return c0 + c1 * x


def _horner(x, cc):
    """polyval :47-51."""
    acc = cc[-1]
    for i in range(2, len(cc) + 1):
        acc = cc[-i] + acc * x
    return acc


def _strip_trailing_zeros(coeff):
    """Polynomial.clean_coeff :155-174."""
    for axis in range(coeff.ndim):
        while coeff.shape[axis] > 1:
            idx = [slice(None)] * coeff.ndim
            idx[axis] = slice(coeff.shape[axis] - 1, coeff.shape[axis])
            if np.sum(np.abs(coeff[tuple(idx)])) != 0:
                break
            coeff = np.delete(coeff, -1, axis=axis)
    return coeff


class _Poly:
    def __init__(self, coeff, clean_zeros=True):
        if isinstance(coeff, list):
            coeff = np.array(coeff)
        if not isinstance(coeff, np.ndarray):
            raise ValueError('Invalid input for Polynomial class object')
        # the reference (:142-146) stores integer tensors uncast; float64 is the superset
        coeff = coeff.astype(np.float64) if coeff.dtype.kind in "iu" else coeff
        self.coeff = _strip_trailing_zeros(coeff) if clean_zeros else coeff
        self.dim = self.coeff.ndim
        self.shape = self.coeff.shape

    def _eval(self, points, rule):
        pts = _as_points(points, self.dim)
        c = self.coeff
        c = rule(pts[:, 0], c.reshape(c.shape + (1,) * pts.ndim))
        for i in range(1, self.dim):
            c = rule(pts[:, i], c)
        return c[0] if len(c) == 1 else c


class MultiCheb(_Poly):
    """Chebyshev-basis coefficient tensor (:250-400)."""

    def __call__(self, points):
        return self._eval(points, _clenshaw)


class MultiPower(_Poly):
    """Power-basis coefficient tensor (:405-640)."""

    def __call__(self, points):
        return self._eval(points, _horner)

    def to_cheb(self):
        """:587-640 -- per axis, x^k expressed in T_j by the x*T_j recurrence."""
        out = self.coeff
        for axis in range(out.ndim):
            moved = np.moveaxis(out, axis, 0)
            acc = np.zeros_like(moved, dtype=np.float64)
            row = np.array([])
            for k, slab in enumerate(moved):
                row = _next_power_row(row)
                acc[:k + 1] += np.einsum("i,...->i...", row, slab)
            out = np.moveaxis(acc, 0, axis)
        return out


def _next_power_row(A):
    """T-coefficients of x^n from those of x^(n-1) (get_new_As :589-614).

    x*T_0 = T_1 and x*T_j = (T_{j+1} + T_{j-1})/2.  Halving is exact, so this
    accumulation is bit-identical to the reference's (A[i-1] + A[i+1])/2.
    """
    n = len(A)
    if n == 0:
        return np.array([1.])
    B = np.zeros(n + 1)
    B[1:] += A / 2          # T_j -> T_{j+1}/2
    B[1] += A[0] / 2        # ... except T_0 -> T_1 in full
    B[:n - 1] += A[1:] / 2  # T_j -> T_{j-1}/2, j >= 1
    return B
