"""Drop-in for yroots.ChebyshevApproximator on a B200.

Same functions and results as /root/reference/yroots/ChebyshevApproximator.py; the heavy parts run
on the device through include/yroots_b200.h:
  * grid samples of MultiCheb / MultiPower objects     -> yr_poly_eval_axis, axis by axis (Clenshaw / Horner)
  * values -> Chebyshev coefficients (dctn type 1, :78-82) -> yr_dct1_axis, axis by axis
  * axis means of |coeff| for the decay tests (:288) and max|values| (:72) -> yr_abs_axis_mean
  * torch-vectorised callables (`device_function(f)`: f takes one float64 tensor per variable and broadcasts)
    are evaluated on the whole grid ON the device -- one call, no host loop, nothing uploaded but the axes
Opaque Python callables are sampled point by point on the host exactly as the reference does (:69) --
an upstream input that bench.py times separately -- and uploaded once.
"""
import itertools
import warnings

import numpy as np

from . import _lib
from .polynomial import MultiCheb, MultiPower
from .runtime import get_engine

macheps = 2.0 ** -52


def transform(x, a, b):
    """[-1,1] -> [a,b] (:8-26)."""
    return ((b - a) * x + (b + a)) / 2


def _is_poly(f):
    return isinstance(f, (MultiCheb, MultiPower))


class device_function:
    """Marks a callable as torch-vectorised: `fn(x0, x1, ...)` accepts float64 torch tensors that broadcast against
    each other and returns their elementwise value (use torch.sin, torch.exp, ... inside).  The approximator then
    samples it on the device in one call per grid (north-star item 1) instead of once per grid point on the host
    (ChebyshevApproximator.py:69).  The object is still an ordinary callable of `dim` scalars, which is how the
    reference's probes (:174-226, :436) and user code call it.

        f = yr.device_function(lambda x, y: torch.sin(x * y) + x * torch.log(y + 3) - x ** 2 + 1 / (y - 4))
        yr.solve([f, g], [-1, -2], [0, 1])
    """
    yr_device = True

    def __init__(self, fn):
        if not callable(fn):
            raise ValueError("Invalid input: input function is not callable")
        self.fn = fn

    def __call__(self, *xs):
        import torch
        if any(isinstance(x, torch.Tensor) for x in xs):
            return self.fn(*xs)
        vals = self.fn(*[torch.as_tensor(np.asarray(x, dtype=np.float64)) for x in xs])
        out = torch.as_tensor(vals, dtype=torch.float64).cpu().numpy()
        return float(out) if out.ndim == 0 else out

    def on_grid(self, axes, device=None):
        """Values on the tensor grid axes[0] x axes[1] x ... as one contiguous float64 tensor on `device`."""
        import torch
        dim = len(axes)
        ts = []
        for d, ax in enumerate(axes):
            shape = [1] * dim
            shape[d] = len(ax)
            ts.append(torch.as_tensor(np.ascontiguousarray(ax, dtype=np.float64)).to(device or "cpu").reshape(shape))
        vals = torch.as_tensor(self.fn(*ts), dtype=torch.float64)
        if device is not None:
            vals = vals.to(device)
        return vals.expand([len(ax) for ax in axes]).contiguous()


class _DeviceGrid:
    """Values of one function on a Chebyshev-Lobatto grid, resident on the device."""

    def __init__(self, eng, f, degs, a, b):
        self.eng, m, lib = eng, eng.mem, eng.lib
        self.npts = [int(d) + 1 for d in degs]
        dim = len(degs)
        axes = [transform(np.cos(np.arange(n + 1) * np.pi / n), lo, hi) for n, lo, hi in zip(degs, a, b)]
        if _is_poly(f):
            # tensor-product evaluation: one axis at a time, same recurrences as Polynomial.__call__
            cur = m.upload(np.ascontiguousarray(f.coeff, dtype=np.float64))
            shape = list(f.coeff.shape)
            for d in range(dim):
                outer = int(np.prod(self.npts[:d], dtype=np.int64))
                inner = int(np.prod(shape[d + 1:], dtype=np.int64))
                x = m.upload(np.ascontiguousarray(axes[d], dtype=np.float64))
                out = m.empty(outer * self.npts[d] * inner * 8)
                _lib.check(lib, lib.yr_poly_eval_axis(cur.ptr, x.ptr, out.ptr, outer, shape[d], self.npts[d], inner,
                                                      f.basis_id, m.stream()), "yr_poly_eval_axis")
                cur = out
            self.values = cur
        elif getattr(f, "yr_device", False):
            # torch-vectorised callable: the grid never exists on the host
            dev = getattr(m, "device", None)
            vals = f.on_grid(axes, dev)
            if dev is not None:
                from .engine import _Buf
                vals = vals.reshape(-1)
                self.values = _Buf(vals.data_ptr(), vals.numel() * 8, vals)
                m.h2d_bytes += sum(len(ax) for ax in axes) * 8
            else:                                                   # CPU test tier: "device memory" is numpy
                self.values = m.upload(vals.numpy())
        else:
            grid = np.meshgrid(*axes, indexing='ij')
            pts = np.column_stack(tuple(g.flatten() for g in grid))
            vals = np.array([f(*p) for p in pts], dtype=np.float64)             # host sampling, as the reference
            self.values = m.upload(vals)
        self.size = int(np.prod(self.npts, dtype=np.int64))

    def sup_norm(self):
        m, lib = self.eng.mem, self.eng.lib
        n0 = self.npts[0]
        mean, mx = m.empty(n0 * 8), m.empty(n0 * 8)
        _lib.check(lib, lib.yr_abs_axis_mean(self.values.ptr, mean.ptr, mx.ptr, 1, n0, self.size // n0, m.stream()),
                   "yr_abs_axis_mean")
        return float(np.max(m.download(mx, n0 * 8).view(np.float64)))

    def coefficients(self):
        """DCT-I along every axis; returns the device buffer of the full (deg+1)^dim coefficient tensor."""
        m, lib = self.eng.mem, self.eng.lib
        cur = self.values
        for d, n in enumerate(self.npts):
            outer = int(np.prod(self.npts[:d], dtype=np.int64))
            inner = int(np.prod(self.npts[d + 1:], dtype=np.int64))
            out = m.empty(self.size * 8)
            _lib.check(lib, lib.yr_dct1_axis(cur.ptr, out.ptr, outer, n, inner, m.stream()), "yr_dct1_axis")
            cur = out
        return cur

    def axis_mean_abs(self, coeff_buf, axis):
        m, lib = self.eng.mem, self.eng.lib
        n = self.npts[axis]
        outer = int(np.prod(self.npts[:axis], dtype=np.int64))
        inner = int(np.prod(self.npts[axis + 1:], dtype=np.int64))
        mean = m.empty(n * 8)
        _lib.check(lib, lib.yr_abs_axis_mean(coeff_buf.ptr, mean.ptr, None, outer, n, inner, m.stream()), "yr_abs_axis_mean")
        return m.download(mean, n * 8).view(np.float64).copy()


def interval_approximate_nd(f, degs, a, b, retSupNorm=False, engine=None):
    """Chebyshev interpolant of f on [a,b] with the given degrees (:28-89)."""
    eng = get_engine() if engine is None else engine
    originalDegs = degs.copy()
    degs[degs == 0] = 1                                # like the reference, this edits the caller's array
    grid = _DeviceGrid(eng, f, degs, a, b)
    supNorm = grid.sup_norm() if retSupNorm else None
    buf = grid.coefficients()
    full = eng.mem.download(buf, grid.size * 8).view(np.float64).reshape(grid.npts)
    coeffs = full[tuple(slice(0, d + 1) for d in originalDegs)]
    return (coeffs, supNorm) if retSupNorm else coeffs


def startedConverging(coeffList, tol):
    """:91-106."""
    return np.all(coeffList[-5:] < tol)


def hasConverged(coeff, coeff2, tol):
    """:108-129."""
    diff = coeff2.copy()
    diff[tuple(slice(0, d) for d in coeff.shape)] -= coeff
    return np.max(np.abs(diff)) < tol


def getFinalDegree(coeff, tol, macheps=2 ** -52):
    """:131-172 -> (degree, epsVal, rho)."""
    settled = int(3 * (len(coeff) - 1) / 4)
    epsVal = 2 * max(macheps, np.max(coeff[settled:]))
    big = np.where(coeff > epsVal)[0]
    degree = 1 if len(big) == 0 else max(1, big[-1])
    if np.all(coeff[1:] < tol):
        degree = 0
    peak = np.argmax(coeff)
    if epsVal == 0:
        epsVal = coeff[peak] * 1e-24
    rho = (coeff[peak] / epsVal) ** (1 / ((degree - peak) + 1))
    return degree, epsVal, rho


_PROBES_1 = [-0.7996847717584993, 0.18546110255464776, -0.13975937255055182, 0., 1., -1.]
_PROBES_2 = [-0.17223860129797386, 0.10828286380141305, -0.5333148248321931, 0.46471703497219596]


def checkConstantInDimension(f, a, b, currDim, relApproxTol, absApproxTol=0):
    """:174-226 -- a dozen point evaluations; stays on the host."""
    call = (lambda p: f(p)) if _is_poly(f) else (lambda p: f(*p))
    dim = len(a)
    for base, probes in ((np.array([0.8984743990614998 ** (k + 1) for k in range(dim)]), _PROBES_1),
                         (np.array([(-0.2598647169391334 * (k + 1) / dim) ** 2 for k in range(dim)]), _PROBES_2)):
        x = transform(base, a, b)
        ref = call(x)
        if np.isclose(ref, 0, rtol=relApproxTol, atol=absApproxTol):
            return False
        for v in transform(np.array(probes), a[currDim], b[currDim]):
            x[currDim] = v
            if not np.allclose(ref, call(x), rtol=relApproxTol, atol=absApproxTol):
                return False
    return True


DECAY_DTYPE = np.dtype([(k, np.float64) for k in ("supnorm", "tol", "started", "converged", "degree", "eps", "rho", "reserved")])


class _Probe:
    """One evaluation of the doubling search: the grid of a degree guess, its coefficient tensor (both on the device) and
    the sub-box of it the reference would keep (:57,:85)."""

    def __init__(self, eng, f, degs, a, b):
        self.keep = [int(d) + 1 for d in degs]                  # extents of the tensor interval_approximate_nd returns
        # The reference turns degree-0 axes into degree 1 IN THE CALLER'S ARRAY (:57), so within one axis' search only the
        # first probe is sliced back to one plane (:85); later probes keep both planes.  Reproduced: the axis means that
        # drive the decisions are averages over exactly the tensors the reference averages over.
        degs[degs == 0] = 1
        self.grid = _DeviceGrid(eng, f, degs, a, b)
        self.coeff = self.grid.coefficients()
        self.shape = list(self.grid.npts)


def _decide(eng, mode, axis, probe, prev, prev_sup, rel, abs_, out_buf, slot):
    """Queues yr_decay_probe for `probe`; its 64-byte answer lands in slot `slot` of out_buf (read later, with the others)."""
    m, lib = eng.mem, eng.lib
    dim = len(probe.shape)
    i64 = lambda v: np.ascontiguousarray(v, dtype=np.int64)
    shape, keep = i64(probe.shape), i64(probe.keep)
    pshape, pkeep = (i64(prev.shape), i64(prev.keep)) if prev is not None else (shape, keep)
    scratch = m.empty((probe.keep[axis] + 128) * 8)
    _lib.check(lib, lib.yr_decay_probe(mode, dim, axis, probe.coeff.ptr, shape.ctypes.data, keep.ctypes.data,
                                       prev.coeff.ptr if prev is not None else None, pshape.ctypes.data, pkeep.ctypes.data,
                                       probe.grid.values.ptr, probe.grid.size, float(prev_sup), float(rel), float(abs_),
                                       scratch.ptr, out_buf.ptr + slot * DECAY_DTYPE.itemsize, m.stream()), "yr_decay_probe")
    probe._scratch = scratch                                     # alive until the answer has been read


def _degree_search(eng, f, a, b, relApproxTol, absApproxTol, out_buf, slot):
    """getChebyshevDegrees (:228-311) as a coroutine: every `yield` is one probe whose decision -- startedConverging,
    hasConverged, getFinalDegree, all evaluated on the device -- the driver hands back as a DECAY_DTYPE record.  Returns
    (degrees, epsilons, rhos)."""
    dim = len(a)
    chebDegrees = [np.inf] * dim
    epsilons, rhos = [], []
    for currDim in range(dim):
        if checkConstantInDimension(f, a, b, currDim, relApproxTol, absApproxTol):
            chebDegrees[currDim] = 0
    for currDim in range(dim):
        if chebDegrees[currDim] == 0:
            epsilons.append(0)
            rhos.append(np.inf)
            continue
        degs = np.array([5] * dim if dim <= 5 else [2] * dim)
        for i in range(dim):
            if chebDegrees[i] < degs[i]:
                degs[i] = chebDegrees[i]
        guess = 8
        while True:
            if guess > 1e5:
                warnings.warn(f"Approximation bound exceeded!\n\nApproximation degree in dimension {currDim} "
                              "has exceeded 1e5, so the process may not finish.\n\nConsider interrupting "
                              "and restarting the process after ensuring that the function(s) inputted are "
                              "continuous and smooth on the approximation interval.\n\n")
            degs[currDim] = guess
            first = _Probe(eng, f, degs, a, b)
            _decide(eng, 0, currDim, first, None, 0.0, relApproxTol, absApproxTol, out_buf, slot)
            ans = yield
            guess *= 2
            if not ans["started"]:
                continue
            degs[currDim] = guess + 1
            second = _Probe(eng, f, degs, a, b)
            _decide(eng, 1, currDim, second, first, ans["supnorm"], relApproxTol, absApproxTol, out_buf, slot)
            ans = yield
            if not ans["converged"]:
                continue
            chebDegrees[currDim] = int(ans["degree"])
            epsilons.append(float(ans["eps"]))
            rhos.append(float(ans["rho"]))
            break
    return np.array(chebDegrees), np.array(epsilons), np.array(rhos)


def _run_searches(eng, jobs, relApproxTol, absApproxTol=0):
    """Runs the degree searches of several (f, a, b) in lock step: per round every unfinished search queues one probe on
    the device and ONE small download brings all their decisions back.  Returns [(degrees, epsilons, rhos)]."""
    m = eng.mem
    out_buf = m.zeros(len(jobs) * DECAY_DTYPE.itemsize)
    gens = [_degree_search(eng, f, a, b, relApproxTol, absApproxTol, out_buf, i) for i, (f, a, b) in enumerate(jobs)]
    results, live = [None] * len(jobs), {}
    for i, g in enumerate(gens):
        try:
            next(g)
            live[i] = g
        except StopIteration as done:
            results[i] = done.value
    while live:
        answers = m.download(out_buf, len(jobs) * DECAY_DTYPE.itemsize).view(DECAY_DTYPE)     # the round's only read-back
        for i in list(live):
            try:
                live[i].send(answers[i].copy())
            except StopIteration as done:
                results[i] = done.value
                del live[i]
    return results


def getChebyshevDegrees(f, a, b, relApproxTol, absApproxTol=0, engine=None):
    """Degree search per axis by doubling (:228-311) -> (degrees, epsilons, rhos)."""
    eng = get_engine() if engine is None else engine
    return _run_searches(eng, [(f, a, b)], relApproxTol, absApproxTol)[0]


def getApproxError(degs, epsilons, rhos):
    """Geometric-tail bound over the 2^dim - 1 truncated index classes (:313-356)."""
    total = 0
    for mask in itertools.product(range(2), repeat=len(degs)):
        if np.sum(mask) == 0:
            continue
        weight, eps = 1, 0
        for i, beyond in enumerate(mask):
            if beyond:
                weight /= (rhos[i] - 1)
                eps = max(eps, epsilons[i])
            else:
                weight *= (degs[i] + 1)
        total += weight * eps
    return total


def _checked_bounds(f, a, b):
    if not hasattr(f, '__call__'):
        raise ValueError("Invalid input: input function is not callable")
    if type(a) == list:
        a = np.array(a)
    if type(b) == list:
        b = np.array(b)
    if type(a) != np.ndarray:
        a = np.array([a])
    if type(b) != np.ndarray:
        b = np.array([b])
    if len(a) != len(b):
        raise ValueError(f"Invalid input: {len(a)} lower bounds were given but {len(b)} upper bounds were given")
    if (b < a).any():
        raise ValueError("Invalid input: at least one lower bound is greater than the corresponding upper bound.")
    mid = [(u + l) / 2 for l, u in zip(a, b)]
    try:
        f(mid) if _is_poly(f) else f(*mid)
    except TypeError:
        raise ValueError("Invalid input: length of the upper/lower bound lists must match the dimension of the function")
    return a, b


def chebApproximate_batch(jobs, relApproxTol=1e-10, engine=None):
    """chebApproximate for a list of (f, a, b): the degree searches advance together (one read-back per round for all of
    them), then every final interpolant is computed.  Returns [(coefficient tensor, error bound)]."""
    eng = get_engine() if engine is None else engine
    jobs = [(f,) + _checked_bounds(f, a, b) for f, a, b in jobs]
    found = _run_searches(eng, jobs, relApproxTol)
    return [(interval_approximate_nd(f, degs, a, b, engine=eng), getApproxError(degs, eps, rhos))
            for (f, a, b), (degs, eps, rhos) in zip(jobs, found)]


def chebApproximate(f, a, b, relApproxTol=1e-10, engine=None):
    """(coefficient tensor, error bound) of f on [a,b] (:358-450)."""
    return chebApproximate_batch([(f, a, b)], relApproxTol, engine)[0]
