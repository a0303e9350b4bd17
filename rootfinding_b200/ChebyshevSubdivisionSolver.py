"""Drop-in for yroots.ChebyshevSubdivisionSolver on a B200.

Same names, argument meaning and error behaviour as
/root/reference/yroots/ChebyshevSubdivisionSolver.py; the work runs in the CUDA kernels
behind include/yroots_b200.h.

  solveChebyshevSubdivision   :1387-1464  one system
  solve_batch                 the same for many independent systems in one device frontier
                              (what the boxes/sec benchmark measures)
  transformChebToInterval     :864-894    batched interval restriction (operator-level entry)
  TransformChebInPlaceND      :339-374
  BoundingIntervalLinearSystem :628-753
  TrackedInterval             :376-576    host-side view of a reported box
"""
import ctypes as C
import warnings

import numpy as np

from . import _lib, postpass
from .runtime import get_engine

macheps = 2.0 ** -52


class SolverOptions:
    """Option bag of the reference (:12-43)."""

    def __init__(self):
        self.verbose = False
        self.exact = False
        self.constant_check = True
        self.low_dim_quadratic_check = True
        self.all_dim_quadratic_check = False
        self.maxZoomCount = 25
        self.level = 0

    def copy(self):
        import copy
        return copy.copy(self)


class TrackedInterval:
    """Host-side interval object with the reference's read API (:376-576).

    Boxes returned by the solver are `postpass.ResultBox` objects (same read API); this class
    exists for user code that builds an interval from an array, e.g.
    tests/test_Combined_Solver.py:269 `TrackedInterval(box).__contains__(root)`.
    """

    def __init__(self, interval):
        self.topInterval = interval
        self.interval = interval
        self.transforms = []
        self.ndim = len(self.interval)
        self.empty = False
        self.finalStep = False
        self.canThrowOutFinalStep = False
        self.possibleDuplicateRoots = []
        self.possibleExtraRoot = False
        self.nextTransformPoints = np.array([0.0394555475981047] * self.ndim)

    def size(self):
        return np.prod(self.interval[:, 1] - self.interval[:, 0])

    def dimSize(self):
        return self.interval[:, 1] - self.interval[:, 0]

    def isPoint(self):
        return bool(np.all(np.abs(self.interval[:, 0] - self.interval[:, 1]) < 1e-32))

    def copy(self):
        twin = TrackedInterval(self.topInterval)
        twin.interval = self.interval.copy()
        twin.nextTransformPoints = self.nextTransformPoints.copy()
        return twin

    def __contains__(self, point):
        return bool(np.all(point >= self.interval[:, 0]) and np.all(point <= self.interval[:, 1]))

    def overlapsWith(self, other):
        a, b = np.asarray(self.interval), np.asarray(other.interval)
        return bool(np.all(a[:, 0] <= b[:, 1]) and np.all(b[:, 0] <= a[:, 1]))

    def __repr__(self):
        return str(self.interval)


def _check_system(Ms, errors):
    n = len(Ms)
    if any(np.ndim(M) != n for M in Ms):
        raise ValueError("Solver Takes in N polynomials of dimension N!")
    if len(Ms) != len(errors):
        raise ValueError("Ms and errors must be same length!")


def _is_out_of_memory(exc):
    """Device out of memory, by TYPE: the library reports it through yr_last_error_code (-> _lib.DeviceOutOfMemory), torch's
    allocator through torch.OutOfMemoryError."""
    if isinstance(exc, _lib.DeviceOutOfMemory):
        return True
    try:
        import torch
        return isinstance(exc, torch.OutOfMemoryError)
    except Exception:
        return False


# Device memory a frontier holds at its widest, as a multiple of the bytes of the system's input coefficients.  Measured
# high-water marks (scripts/profile_solve.py, library pools included): 2-D degree 20 / 50 / 100: 830 / 1050 / 1150 x,
# 3-D degree 10: 6200 x, 4-D degree 4 / 6: 5500 / 25000 x (profiles/r02_configs.txt).  The table is deliberately on the
# safe side; solve_batch uses it to cut a batch into sub-batches that fit BEFORE launching anything.
_PEAK_FACTOR = {1: 400, 2: 1500, 3: 8000, 4: 30000, 5: 30000, 6: 30000}


def _batch_ranges_for_memory(sizes, dim, eng):
    """Consecutive sub-batches whose estimated peak frontier fits the free device memory (None: one batch is fine)."""
    info = getattr(eng.mem, "free_bytes", None)
    if info is None:
        return None
    budget = 0.6 * info()
    need = sizes.astype(np.float64) * 8.0 * _PEAK_FACTOR.get(dim, 30000)
    if need.sum() <= budget or len(sizes) < 2:
        return None
    ranges, lo, acc = [], 0, 0.0
    for i, b in enumerate(need):
        if acc + b > budget and i > lo:
            ranges.append((lo, i))
            lo, acc = i, 0.0
        acc += b
    ranges.append((lo, len(sizes)))
    return ranges if len(ranges) > 1 else None


def _solve_chunks(systems, errors, ranges, returnBoundingBoxes, exact, constant_check, low_dim_quadratic_check,
                  all_dim_quadratic_check, use_fma, eng, return_session, trim_policy=0):
    """solve_batch over consecutive sub-batches, one after the other; the last session carries all of them in
    `chunk_sessions` (statistics)."""
    roots, boxes, sessions = [], [], []
    for lo, hi in ranges:
        out, sess = solve_batch(systems[lo:hi], errors[lo:hi], returnBoundingBoxes, exact, constant_check,
                                low_dim_quadratic_check, all_dim_quadratic_check, use_fma, eng, True, trim_policy)
        if returnBoundingBoxes:
            roots += out[0]; boxes += out[1]
        else:
            roots += out
        if sess is not None:
            chunk = getattr(sess, "chunk_sessions", [sess])
            if not return_session:
                for c in chunk:                        # the next chunk needs the device memory this one still pins
                    c.release_device()
            sessions += chunk
        del sess
    last = sessions[-1] if sessions else None
    if last is not None:
        last.chunk_sessions = sessions
    out = (roots, boxes) if returnBoundingBoxes else roots
    return (out, last) if return_session else out


def solve_batch(systems, errors, returnBoundingBoxes=False, exact=False, constant_check=True,
                low_dim_quadratic_check=True, all_dim_quadratic_check=False, use_fma=False, engine=None,
                return_session=False, trim_policy=0):
    """Solve many independent systems (each a list of `dim` tensors) in one device frontier.

    `trim_policy` (not a parameter of the reference): 0 = the reference's trimMs (:1131-1171); 1, 2, 3 = trimMsOptimized1/2/3 of
    its experiment file tests/TrimMs optimization/ChebyshevSubdivisionSolverWithOptimizedTrimMs.py:1167-1271, plugged into
    the same call site (:1232).  Policies 1-3 run on the level pipeline.

    Returns a list with one root array per system (and a list of box lists when returnBoundingBoxes).  Roots of a system
    come in depth-first order of the boxes that produced them (re-solved merged boxes at their job's position).  The
    reference appends a call's exterior results after its interior ones at every level of the recursion
    (ChebyshevSubdivisionSolver.py:1318-1385), so its order differs; the root SET is the reference's and both test suites
    compare sets (tests/test_Combined_Solver.py:12-29 of the reference).
    """
    if trim_policy not in (0, 1, 2, 3):
        raise ValueError("trim_policy must be 0 (trimMs) or 1, 2, 3 (trimMsOptimized1/2/3)")
    for Ms, e in zip(systems, errors):
        _check_system(Ms, e)
    if len(systems) == 0:                                  # an empty batch has an empty answer and needs no device
        out = ([], []) if returnBoundingBoxes else []
        return (out, None) if return_session else out
    eng = get_engine() if engine is None else engine
    # Very large batches are solved as a sequence of sub-batches sized for the device: box records of a 2-D solve take
    # ~1 KB per box and a random system produces ~1.3 boxes per coefficient, so 24 M coefficients (4096 degree-50
    # systems) keep the frontier's tables well inside 180 GB, and a sub-batch that size already fills the machine.
    limit = getattr(eng, "max_batch_coeffs", 24_000_000)
    if len(systems) > 1:
        sizes = np.fromiter((sum(np.size(M) for M in Ms) for Ms in systems), dtype=np.int64, count=len(systems))
        total = int(sizes.sum())
        if total > limit:
            nchunk = min(len(systems), -(-total // limit))
            edges = np.searchsorted(np.cumsum(sizes), np.arange(1, nchunk) * (total / nchunk), side="left") + 1
            bounds = np.unique(np.clip(np.r_[0, edges, len(systems)], 0, len(systems)))
            if len(bounds) > 2:
                return _solve_chunks(systems, errors, list(zip(bounds[:-1], bounds[1:])), returnBoundingBoxes, exact,
                                     constant_check, low_dim_quadratic_check, all_dim_quadratic_check, use_fma, eng, return_session, trim_policy)
        ranges = _batch_ranges_for_memory(sizes, len(systems[0]), eng)
        if ranges is not None:
            return _solve_chunks(systems, errors, ranges, returnBoundingBoxes, exact, constant_check, low_dim_quadratic_check,
                                 all_dim_quadratic_check, use_fma, eng, return_session, trim_policy)
    sess = None
    try:
        sess = eng.session(systems, errors,
                           exact=int(bool(exact)), constant_check=int(bool(constant_check)),
                           low_dim_quad_check=int(bool(low_dim_quadratic_check)),
                           all_dim_quad_check=int(bool(all_dim_quadratic_check)), use_fma=int(bool(use_fma)),
                           trim_policy=int(trim_policy))
        sess.run_jobs(sess.unit_jobs())
    except RuntimeError as exc:
        if not _is_out_of_memory(exc) or len(systems) < 2:
            raise
        sess = None                                        # (the retry runs outside this block: the traceback pins
                                                           #  the failed attempt's device buffers until it is gone)
    if sess is None:
        # the frontier of the whole batch does not fit the device: solve it in two halves, one after the other
        # (systems never interact, so the answer is the same; a single system that does not fit still raises)
        eng.release_cached()
        half = len(systems) // 2
        return _solve_chunks(systems, errors, [(0, half), (half, len(systems))], returnBoundingBoxes, exact, constant_check,
                             low_dim_quadratic_check, all_dim_quadratic_check, use_fma, eng, return_session, trim_policy)
    res, keep, dup, extra = postpass.assemble(sess, need_records=returnBoundingBoxes)
    if sess.stats.max_level >= 25:
        warnings.warn("Extreme subdivision depth!\nSubdivision on the search interval has now reached"
                      " at least depth 25, which is unusual. The solver may not finish running."
                      "Ensure the input functions meet the requirements of being continuous, smooth,"
                      "and having only finitely many simple roots on the search interval.")
    elif sess.stats.max_level >= 15:
        warnings.warn("High subdivision depth!\nSubdivision on the search interval has now reached"
                      " at least depth 15. Runtime may be prolonged.")
    roots, boxes = postpass.per_system_roots(res, keep, dup, extra, sess.nsys, sess.dim, want_boxes=returnBoundingBoxes,
                                             sorted_fields=getattr(sess, "sorted_fields", None),
                                             late_by_system=getattr(sess, "late_by_system", None),
                                             redo_systems=getattr(sess, "redo_systems", None), alive=getattr(sess, "alive", None))
    out = (roots, boxes) if returnBoundingBoxes else roots
    return (out, sess) if return_session else out


def solveChebyshevSubdivision(Ms, errors, verbose=False, returnBoundingBoxes=False, exact=False,
                              constant_check=True, low_dim_quadratic_check=True, all_dim_quadratic_check=False, trim_policy=0):
    """Roots (and optionally bounding boxes) of one Chebyshev system on [-1,1]^n (:1387-1464).  `trim_policy`: see solve_batch."""
    _check_system(Ms, errors)
    if verbose:
        print("Finding roots...", end=' ')
    out = solve_batch([Ms], [np.asarray(errors, dtype=np.float64)], True, exact, constant_check,
                      low_dim_quadratic_check, all_dim_quadratic_check, trim_policy=trim_policy)
    roots, boxes = out[0][0], out[1][0]
    dim = len(Ms)
    roots = np.array(roots).reshape(-1, dim) if len(roots) else np.array([])
    if verbose:
        finish_string = '\n' + f"Found {len(roots)} roots"
        print((finish_string if len(roots) != 1 else finish_string[:-1]), end='\n\n')
    if returnBoundingBoxes:
        return roots, boxes
    return roots


# ---- operator-level entry points -----------------------------------------------------------
def restrict_systems(systems, errors, alphas, betas, exact=False, use_fma=False, engine=None):
    """Batched transformChebToInterval: system s restricted to x -> alphas[s]*x + betas[s].

    Returns (list of lists of tensors, errors [nsys][dim]).
    """
    eng = get_engine() if engine is None else engine
    sess = eng.session(systems, errors, exact=int(bool(exact)), use_fma=int(bool(use_fma)))
    D, n = sess.dim, sess.nsys
    jobs = sess.unit_jobs()
    jobs["restrict"] = np.ones(n, np.int32)
    jobs["alpha"] = np.asarray(alphas, dtype=np.float64).reshape(n, D)
    jobs["beta"] = np.asarray(betas, dtype=np.float64).reshape(n, D)
    recs, data, c = sess._init_boxes(jobs)
    bdt = _lib.box_dtype(sess.lib, D)
    boxes = sess.mem.download(recs, n * bdt.itemsize).view(bdt)
    flat = sess.mem.download(data, int(c["data_used"]) * 8).view(np.float64)
    out = [None] * n
    errs = np.zeros((n, D))
    for b in boxes:
        s = int(b["system"])
        off = int(b["data_off"])
        Ms = []
        for p in range(D):
            shp = tuple(int(v) for v in b["shape"][p])
            sz = int(np.prod(shp))
            Ms.append(flat[off:off + sz].reshape(shp).copy())
            off += sz
        out[s] = Ms
        errs[s] = b["err"]
    return out, errs


def transformChebToInterval(Ms, alphas, betas, errors, exact):
    """:864-894 for one system."""
    out, errs = restrict_systems([Ms], [errors], [alphas], [betas], exact)
    return out[0], errs[0]


def TransformChebInPlaceND(coeffs, dim, alpha, beta, exact):
    """:339-374 -- one axis of one tensor (the tensor is embedded as polynomial 0 of a system)."""
    coeffs = np.asarray(coeffs, dtype=np.float64)
    if (alpha == 1.0 and beta == 0.0) or coeffs.shape[dim] == 1:
        return coeffs
    D = coeffs.ndim
    al, be = np.ones(D), np.zeros(D)
    al[dim], be[dim] = alpha, beta
    filler = [np.zeros((1,) * D) for _ in range(D - 1)]
    out, _ = restrict_systems([[coeffs] + filler], [np.zeros(D)], [al], [be], exact)
    return out[0][0]


def getLinearTerms(M):
    """:578-605."""
    M = np.asarray(M)
    out = []
    for d in range(M.ndim):
        idx = [0] * M.ndim
        idx[d] = 1
        out.append(0 if M.shape[d] == 1 else M[tuple(idx)])
    return out


def bils_batched(systems, errors, final_steps, engine=None):
    """BoundingIntervalLinearSystem for many systems: (intervals [n][dim][2], flags [n][3])."""
    eng = get_engine() if engine is None else engine
    n, D = len(systems), len(systems[0])
    lin = np.array([[getLinearTerms(M) for M in Ms] for Ms in systems], dtype=np.float64)
    c0 = np.array([[np.asarray(M).ravel()[0] for M in Ms] for Ms in systems], dtype=np.float64)
    sumabs = np.array([[np.sum(np.abs(M)) for M in Ms] for Ms in systems], dtype=np.float64)
    m = eng.mem
    d_iv, d_fl = m.empty(n * D * 16), m.empty(n * 12)
    bufs = [m.upload(lin), m.upload(c0), m.upload(sumabs), m.upload(np.asarray(errors, dtype=np.float64).reshape(n, D)),
            m.upload(np.asarray(final_steps, dtype=np.int32))]
    _lib.check(eng.lib, eng.lib.yr_bils_batched(D, n, bufs[0].ptr, bufs[1].ptr, bufs[2].ptr, bufs[3].ptr, bufs[4].ptr,
                                                d_iv.ptr, d_fl.ptr, m.stream()), "yr_bils_batched")
    iv = m.download(d_iv, n * D * 16).view(np.float64).reshape(n, D, 2)
    fl = m.download(d_fl, n * 12).view(np.int32).reshape(n, 3).astype(bool)
    return iv, fl


def BoundingIntervalLinearSystem(Ms, errors, finalStep, macheps=2 ** -52):
    """:628-753 -> (interval, changed, should_stop, throwOut)."""
    iv, fl = bils_batched([Ms], [errors], [int(bool(finalStep))])
    return iv[0], bool(fl[0][0]), bool(fl[0][1]), bool(fl[0][2])


def frontier_tensor_ops(op, systems, errors, alphas=None, betas=None, intervals=None, next_mids=None, levels=None,
                        flat=True, use_fma=False, engine=None):
    """The 2-D frontier pipeline's tensor kernels on caller-supplied boxes (include/yroots_b200.h: yr_f2_tensor_ops):
    op = "restrict" -- transformChebToInterval (:864-894) with alphas / betas [n][2];
    op = "subdivide" -- getSubdivisionIntervals (:1051-1129) with the boxes' tracked `intervals` [n][2][2], `next_mids` [n][2]
    and `levels` [n].  Returns one dict per output box: Ms (two tensors), errors, sumabs, trim_shapes, trim_errors
    (what trimMs :1131-1171 would leave), interval, next_mid, parent (index of the input box)."""
    eng = get_engine() if engine is None else engine
    n = len(systems)
    code = {"restrict": 1, "subdivide": 2}[op]
    shape = np.array([[M.shape for M in Ms] for Ms in systems], dtype=np.int32).reshape(n, 2, 2)
    sizes = shape.prod(axis=2).reshape(-1).astype(np.uint64)
    off = np.ascontiguousarray(np.cumsum(sizes) - sizes, dtype=np.uint64)
    coeffs = np.concatenate([np.ascontiguousarray(M, dtype=np.float64).reshape(-1) for Ms in systems for M in Ms])
    err = np.ascontiguousarray(errors, dtype=np.float64).reshape(n, 2)
    d = _lib.YrF2OpsDesc()
    d.op, d.flat, d.n = code, int(bool(flat)), n
    keep = [coeffs, off, shape, err]
    d.coeffs, d.off, d.shape, d.err = coeffs.ctypes.data, off.ctypes.data, shape.ctypes.data, err.ctypes.data
    for name, arr, dt, shp in (("alpha", alphas, np.float64, (n, 2)), ("beta", betas, np.float64, (n, 2)),
                               ("iv", intervals, np.float64, (n, 2, 2)), ("next_mid", next_mids, np.float64, (n, 2)),
                               ("level", levels, np.int32, (n,))):
        if arr is not None:
            a = np.ascontiguousarray(arr, dtype=dt).reshape(shp)
            keep.append(a)
            setattr(d, name, a.ctypes.data)
    d.prm = eng.params(use_fma=int(bool(use_fma)))
    cap = n * (4 if code == 2 else 1)
    total = int(sizes.sum()) * (4 if code == 2 else 1) + 16
    n_out = np.zeros(1, dtype=np.uint64)
    out = dict(coeffs=np.zeros(total), off=np.zeros((cap, 2), np.uint64), shape=np.zeros((cap, 2, 2), np.int32), err=np.zeros((cap, 2)),
               sumabs=np.zeros((cap, 2)), tshape=np.zeros((cap, 2, 2), np.int32), terr=np.zeros((cap, 2)), iv=np.zeros((cap, 2, 2)),
               mid=np.zeros((cap, 2)), parent=np.zeros(cap, np.int32))
    d.cap_out, d.cap_coeffs, d.n_out = cap, total, n_out.ctypes.data
    d.out_coeffs, d.out_off, d.out_shape, d.out_err = (out[k].ctypes.data for k in ("coeffs", "off", "shape", "err"))
    d.out_sumabs, d.out_trim_shape, d.out_trim_err = (out[k].ctypes.data for k in ("sumabs", "tshape", "terr"))
    d.out_iv, d.out_next_mid, d.out_parent = (out[k].ctypes.data for k in ("iv", "mid", "parent"))
    import ctypes as C
    _lib.check(eng.lib, eng.lib.yr_f2_tensor_ops(C.byref(d), eng.mem.stream()), "yr_f2_tensor_ops")
    boxes = []
    for q in range(int(n_out[0])):
        Ms = []
        for p in range(2):
            shp = tuple(int(v) for v in out["shape"][q, p])
            o = int(out["off"][q, p])
            Ms.append(out["coeffs"][o:o + shp[0] * shp[1]].reshape(shp).copy())
        boxes.append(dict(Ms=Ms, errors=out["err"][q].copy(), sumabs=out["sumabs"][q].copy(), trim_shapes=out["tshape"][q].copy(),
                          trim_errors=out["terr"][q].copy(), interval=out["iv"][q].copy(), next_mid=out["mid"][q].copy(),
                          parent=int(out["parent"][q])))
    return boxes
