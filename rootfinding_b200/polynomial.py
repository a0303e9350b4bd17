"""Drop-in for yroots.polynomial: MultiCheb / MultiPower coefficient-tensor objects.

The objects are host-side (numpy) like the reference's (/root/reference/yroots/polynomial.py:95-640);
what changes is where their *grid* evaluation happens when the approximator samples them: on the
device, axis by axis (ChebyshevApproximator.interval_approximate_nd -> yr_poly_eval_axis), with the
same Clenshaw / Horner recurrences and rounding as `__call__` below.
"""
import numpy as np


def _points(points, dim):
    """Polynomial.__call__ :176-205."""
    pts = np.array(points)
    if pts.ndim == 0:
        pts = np.array([pts])
    if pts.ndim == 1:
        pts = pts.reshape(1, pts.shape[0]) if dim > 1 else pts.reshape(pts.shape[0], 1)
    if pts.shape[1] != dim:
        raise ValueError('Dimension of points does not match dimension of polynomial!')
    return pts


def chebval(x, cc):
    """Clenshaw along the leading axis (:59-74)."""
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


def polyval(x, cc):
    """Horner along the leading axis (:47-51)."""
    acc = cc[-1]
    for i in range(2, len(cc) + 1):
        acc = cc[-i] + acc * x
    return acc


def match_size(a, b):
    shape = np.maximum(a.shape, b.shape)
    A, B = np.zeros(shape), np.zeros(shape)
    A[tuple(slice(0, n) for n in a.shape)] = a
    B[tuple(slice(0, n) for n in b.shape)] = b
    return A, B


class Polynomial(object):
    """Common part of MultiCheb and MultiPower (:95-245)."""
    _rule = None
    basis_id = -1

    def __init__(self, coeff, clean_zeros=True):
        if isinstance(coeff, list):
            coeff = np.array(coeff)
        if not isinstance(coeff, np.ndarray):
            raise ValueError('Invalid input for Polynomial class object')
        if coeff.dtype.kind in "iub":
            coeff = coeff.astype(np.float64)       # the reference leaves integer tensors uncast (:142-146) and then fails
        self.coeff = coeff
        if clean_zeros:
            self.clean_coeff()
        self.dim = self.coeff.ndim
        self.shape = self.coeff.shape
        self.jac = None

    def clean_coeff(self):
        """Strip all-zero trailing hyperplanes (:155-174)."""
        for axis in range(self.coeff.ndim):
            while self.coeff.shape[axis] > 1:
                idx = [slice(None)] * self.coeff.ndim
                idx[axis] = slice(self.coeff.shape[axis] - 1, self.coeff.shape[axis])
                if np.sum(abs(self.coeff[tuple(idx)])) != 0:
                    break
                self.coeff = np.delete(self.coeff, -1, axis=axis)

    def __call__(self, points):
        pts = _points(points, self.dim)
        c = self.coeff
        c = type(self)._rule(pts[:, 0], c.reshape(c.shape + (1,) * pts.ndim))
        for i in range(1, self.dim):
            c = type(self)._rule(pts[:, i], c)
        return c[0] if len(c) == 1 else c

    def evaluate_grid(self, xyz):
        """Values on the tensor grid spanned by the columns of xyz (:344-371, :530-557)."""
        xyz = _points(xyz, self.dim)
        c = self.coeff
        for i in range(xyz.shape[1]):
            c = type(self)._rule(xyz[:, i], c.reshape(c.shape + (1,) * xyz[:, i].ndim))
        return c[0] if np.prod(c.shape) == 1 else c

    def __eq__(self, other):
        return self.shape == other.shape and bool(np.allclose(self.coeff, other.coeff))

    def __ne__(self, other):
        return not (self == other)

    def __repr__(self):
        return str(self.coeff)

    __str__ = __repr__
    __hash__ = None

    def _binary(self, other, op, **kw):
        a, b = (self.coeff, other.coeff) if self.shape == other.shape else match_size(self.coeff, other.coeff)
        return type(self)(op(a, b), **kw)


class MultiCheb(Polynomial):
    """Chebyshev-basis tensor: entry (i,j,...) multiplies T_i(x) T_j(y) ... (:250-400)."""
    _rule = staticmethod(chebval)
    basis_id = 0

    def __add__(self, other):
        return self._binary(other, np.add)

    def __sub__(self, other):
        return self._binary(other, np.subtract, clean_zeros=False)

    def grad(self, point):
        from numpy.polynomial import chebyshev as cheb
        if len(point) != self.dim:
            raise ValueError('Cannot evaluate polynomial in {} variables at point {}'.format(self.dim, point))
        if self.jac is None:
            self.jac = [cheb.chebder(self.coeff, axis=i) for i in range(self.dim)]
        return np.array([MultiCheb(j, clean_zeros=False)(point) for j in self.jac], dtype=complex)


class MultiPower(Polynomial):
    """Power-basis tensor: entry (i,j,...) multiplies x^i y^j ... (:405-640)."""
    _rule = staticmethod(polyval)
    basis_id = 1

    def __add__(self, other):
        return self._binary(other, np.add, clean_zeros=False)

    def __sub__(self, other):
        return self._binary(other, np.subtract, clean_zeros=False)

    def __mul__(self, other):
        from scipy.signal import convolve
        a, b = (self.coeff, other.coeff) if self.shape == other.shape else match_size(self.coeff, other.coeff)
        return MultiPower(convolve(a, b))

    def grad(self, point):
        from numpy.polynomial import polynomial as poly
        if len(point) != self.dim:
            raise ValueError('Cannot evaluate polynomial in {} variables at point {}'.format(self.dim, point))
        if self.jac is None:
            self.jac = [poly.polyder(self.coeff, axis=i) for i in range(self.dim)]
        return np.array([MultiPower(j, clean_zeros=False)(point) for j in self.jac], dtype=complex)

    def to_cheb(self, engine=None):
        """Chebyshev coefficient tensor of the same polynomial (:587-640): per axis, the T-expansion
        of x^k is built from that of x^(k-1) by x T_0 = T_1, x T_j = (T_{j+1} + T_{j-1}) / 2.

        With an `engine` the tensor is transformed on the device (yr_upper_axis_apply, one launch per axis; the small
        power -> Chebyshev matrices are built here with the same recurrence): same operations in the same order, so the
        result is bit-identical to the host loop."""
        if engine is not None:
            return _to_cheb_device(np.ascontiguousarray(self.coeff, dtype=np.float64), engine)
        out = self.coeff
        for axis in range(out.ndim):
            moved = np.moveaxis(out, axis, 0)
            acc = np.zeros_like(moved, dtype=np.float64)
            row = np.array([])
            for k, slab in enumerate(moved):
                row = _times_x(row)
                acc[:k + 1] += np.einsum("i,...->i...", row, slab)
            out = np.moveaxis(acc, 0, axis)
        return out


def _power_to_cheb_matrix(n):
    """P[i][k] = coefficient of T_i in x^k, k < n (upper triangular), from the recurrence of _times_x."""
    P = np.zeros((n, n))
    row = np.array([])
    for k in range(n):
        row = _times_x(row)
        P[:k + 1, k] = row
    return P


def _to_cheb_device(coeff, eng):
    from . import _lib
    m, lib = eng.mem, eng.lib
    shape = coeff.shape
    cur = m.upload(coeff)
    for axis, n in enumerate(shape):
        outer = int(np.prod(shape[:axis], dtype=np.int64))
        inner = int(np.prod(shape[axis + 1:], dtype=np.int64))
        P = m.upload(_power_to_cheb_matrix(n))
        out = m.empty(coeff.size * 8)
        _lib.check(lib, lib.yr_upper_axis_apply(cur.ptr, P.ptr, out.ptr, outer, n, inner, m.stream()), "yr_upper_axis_apply")
        cur = out
    return m.download(cur, coeff.size * 8).view(np.float64).reshape(shape).copy()


def _times_x(A):
    n = len(A)
    if n == 0:
        return np.array([1.])
    B = np.zeros(n + 1)
    B[1:] += A / 2
    B[1] += A[0] / 2
    B[:n - 1] += A[1:] / 2
    return B
