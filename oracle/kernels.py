"""ORACLE -- test infrastructure only (see oracle/README.md).

ctypes front-end for oracle/cheb_kernels.c, the plain-C restatement of the
reference's numba kernels (ChebyshevSubdivisionSolver.py:45-337, :608-626).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_kernels.so")
_lib = None


def build(force=False):
    """Compile cheb_kernels.c with gcc (make -C oracle)."""
    if force or not os.path.exists(_SO) or (
            os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "cheb_kernels.c"))):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        for name in ("orc_restrict_axis0", "orc_restrict_axis0_exact"):
            f = getattr(L, name)
            f.restype = ctypes.c_long
            f.argtypes = [dp, ctypes.c_long, ctypes.c_long, ctypes.c_double, ctypes.c_double, dp]
        L.orc_restrict_axis0_split.restype = ctypes.c_long
        L.orc_restrict_axis0_split.argtypes = [dp, ctypes.c_long, ctypes.c_long, ctypes.c_double, dp]
        L.orc_linear_check.restype = None
        L.orc_linear_check.argtypes = [dp, dp, dp, ctypes.c_long, dp, dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def restrict_axis(coeffs, axis, alpha, beta, exact=False):
    """M' = C(alpha,beta) applied along `axis` (TransformChebInPlaceND, :339-374).

    Returns a new array whose extent along `axis` is the
    reference's `maxRow`; identity transforms and extent-1 axes pass through.
    """
    coeffs = np.asarray(coeffs, dtype=np.float64)
    if (alpha == 1.0 and beta == 0.0) or coeffs.shape[axis] == 1:
        return coeffs
    view = np.moveaxis(coeffs, axis, 0)
    # numba allocates its result like the array it is handed: Fortran order for
    # an F-contiguous argument (the 2-D transposed case), C order otherwise.
    # The values do not depend on it; later np.sum(|M|) rounding does.
    f_layout = view.flags.f_contiguous and not view.flags.c_contiguous
    moved = np.ascontiguousarray(view)
    n = moved.shape[0]
    m = moved.size // n
    out = np.empty_like(moved)
    fn = lib().orc_restrict_axis0_exact if exact else lib().orc_restrict_axis0
    rows = fn(_dp(moved), n, m, float(alpha), float(beta), _dp(out))
    if f_layout:
        out = np.asfortranarray(out)
    return np.moveaxis(out[:rows], 0, axis)


def linear_check(total_errs, A, consts):
    """linearCheck1 (:608-626): per-column bounds from every row's linear term."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    d = A.shape[0]
    t = np.ascontiguousarray(total_errs, dtype=np.float64)
    c = np.ascontiguousarray(consts, dtype=np.float64)
    lo = np.empty(d)
    hi = np.empty(d)
    lib().orc_linear_check(_dp(t), _dp(A), _dp(c), d, _dp(lo), _dp(hi))
    return lo, hi
