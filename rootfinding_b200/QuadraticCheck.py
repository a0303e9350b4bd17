"""Drop-in for yroots.QuadraticCheck.quadratic_check (QuadraticCheck.py:25-31) on the device."""
import numpy as np

from . import _lib
from .runtime import get_engine


def quadratic_check_batched(tensors, tols, nd_check=False, engine=None):
    """True where the quadratic bound proves there is no root. All tensors share one dimension."""
    eng = get_engine() if engine is None else engine
    n, D = len(tensors), np.ndim(tensors[0])
    flat = [np.ascontiguousarray(t, dtype=np.float64).reshape(-1) for t in tensors]
    off = np.zeros(n, dtype=np.int64)
    off[1:] = np.cumsum([f.size for f in flat])[:-1]
    shapes = np.array([np.shape(t) for t in tensors], dtype=np.int32).reshape(n, D)
    m = eng.mem
    bufs = [m.upload(np.concatenate(flat)), m.upload(off), m.upload(shapes), m.upload(np.asarray(tols, dtype=np.float64))]
    d_out = m.empty(n * 4)
    _lib.check(eng.lib, eng.lib.yr_quadcheck_batched(D, n, bufs[0].ptr, bufs[1].ptr, bufs[2].ptr, bufs[3].ptr,
                                                     int(bool(nd_check)), d_out.ptr, m.stream()), "yr_quadcheck_batched")
    return m.download(d_out, n * 4).view(np.int32).astype(bool)


def quadratic_check(test_coeff, tol, nd_check=False):
    return bool(quadratic_check_batched([test_coeff], [tol], nd_check)[0])
