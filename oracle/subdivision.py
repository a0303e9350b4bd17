"""ORACLE -- test infrastructure only (see oracle/README.md).

CPU restatement (numpy + oracle/cheb_kernels.c) of the reference's Chebyshev
subdivision solver, /root/reference/yroots/ChebyshevSubdivisionSolver.py.
Each function cites the reference lines it follows.  The recursion tree, the
order of floating-point operations that numpy makes visible (np.sum,
np.linalg.svd, elementwise expressions) and every threshold are kept so that
this oracle can be pinned against the reference itself (oracle/gen_golden.py,
tests/test_oracle_vs_golden.py).

Product code never imports this module.
"""
import itertools
import warnings

import numpy as np

from . import kernels
from .quadratic import quadratic_check

MACHEPS = 2.0 ** -52
FIRST_SPLIT = 0.0394555475981047          # :414


# ----------------------------------------------------------------------------
# error-free transformations (:755-805), numpy-vectorised like the *_NoNumba ones
# ----------------------------------------------------------------------------
def two_sum(a, b):
    x = a + b
    z = x - a
    return x, (a - (x - z)) + (b - z)


def veltkamp_split(a):
    c = (2 ** 27 + 1) * a
    x = c - (c - a)
    return x, a - x


def two_prod(a, b):
    x = a * b
    a1, a2 = veltkamp_split(a)
    b1, b2 = veltkamp_split(b)
    return x, a2 * b2 - (((x - a1 * b1) - a2 * b1) - a1 * b2)


# ----------------------------------------------------------------------------
# per-box interval state (TrackedInterval, :376-576)
# ----------------------------------------------------------------------------
class Box:
    """State that travels with one subdivision box.

    interval  : [dim,2] current box in the coordinates of the solve's unit cube,
                updated in plain doubles, exactly at +-1 (:445-454)
    maps      : the (alpha[dim], beta[dim]) restrictions applied so far; replayed
                with compensated arithmetic for the reported box / root (:460-514)
    """

    def __init__(self, interval):
        self.topInterval = interval
        self.interval = interval
        self.transforms = []
        self.ndim = len(interval)
        self.empty = False
        self.finalStep = False
        self.canThrowOutFinalStep = False
        self.possibleDuplicateRoots = []
        self.possibleExtraRoot = False
        self.nextTransformPoints = np.array([FIRST_SPLIT] * self.ndim)

    # -- :416-418
    def canThrowOut(self):
        return (not self.finalStep) or self.canThrowOutFinalStep

    # -- :420-454
    def addTransform(self, sub):
        inverted = np.any(sub[:, 0] > sub[:, 1])
        if inverted and self.canThrowOut():
            self.empty = True
            return
        if inverted:
            lo = np.maximum(np.minimum(sub[:, 0], 1.0), -1.0)
            hi = np.maximum(np.minimum(sub[:, 1], 1.0), lo)
            sub[:, 0] = lo
            sub[:, 1] = hi
        a1, b1 = sub.T
        a2, b2 = self.interval.T
        alpha1, beta1 = (b1 - a1) / 2, (b1 + a1) / 2
        alpha2, beta2 = (b2 - a2) / 2, (b2 + a2) / 2
        self.transforms.append(np.array([alpha1, beta1]))
        for d in range(self.ndim):
            for side in range(2):
                x = sub[d][side]
                if x == -1.0:
                    self.interval[d][side] = self.interval[d][0]
                elif x == 1.0:
                    self.interval[d][side] = self.interval[d][1]
                else:
                    self.interval[d][side] = alpha2[d] * x + beta2[d]

    def getLastTransform(self):
        return self.transforms[-1]

    def _replay(self, maps):
        """Push [-1,1]^dim through the recorded maps, innermost first (:473-483)."""
        val = self.topInterval.T
        err = np.zeros_like(val)
        for alpha, beta in maps[::-1]:
            val, e = two_prod(val, alpha)
            err = alpha * err + e
            val, e = two_sum(val, beta)
            err += e
        return val.T, err.T

    # -- :460-489
    def getFinalInterval(self):
        maps = self.preFinalTransforms if self.finalStep else self.transforms
        val, err = self._replay(maps)
        self.finalInterval = val + err
        self.finalAlpha, ea = two_sum(-val[:, 0] / 2, val[:, 1] / 2)
        self.finalAlpha += ea + (err[:, 1] - err[:, 0]) / 2
        self.finalBeta, eb = two_sum(val[:, 0] / 2, val[:, 1] / 2)
        self.finalBeta += eb + (err[:, 1] + err[:, 0]) / 2
        return self.finalInterval

    # -- :491-514
    def getFinalPoint(self):
        if not self.finalStep:
            self.root = (self.finalInterval[:, 0] + self.finalInterval[:, 1]) / 2
        else:
            val, err = self._replay(self.transforms)
            fin = val + err
            self.root = (fin[:, 0] + fin[:, 1]) / 2
        return self.root

    def size(self):
        return np.prod(self.interval[:, 1] - self.interval[:, 0])

    def dimSize(self):
        return self.interval[:, 1] - self.interval[:, 0]

    def finalDimSize(self):
        return self.finalInterval[:, 1] - self.finalInterval[:, 0]

    # -- :528-542
    def copy(self):
        twin = Box(self.topInterval)
        twin.interval = self.interval.copy()
        twin.transforms = self.transforms.copy()
        twin.empty = self.empty
        twin.nextTransformPoints = self.nextTransformPoints.copy()
        if self.finalStep:
            twin.finalStep = True
            twin.canThrowOutFinalStep = self.canThrowOutFinalStep
            twin.possibleDuplicateRoots = self.possibleDuplicateRoots.copy()
            twin.possibleExtraRoot = self.possibleExtraRoot
            twin.preFinalInterval = self.preFinalInterval.copy()
            twin.preFinalTransforms = self.preFinalTransforms.copy()
        return twin

    def __contains__(self, point):
        return np.all(point >= self.interval[:, 0]) and np.all(point <= self.interval[:, 1])

    # -- :548-556
    def overlapsWith(self, other):
        for (a1, b1), (a2, b2) in zip(self.getIntervalForCombining(), other.getIntervalForCombining()):
            if a1 > b2 or a2 > b1:
                return False
        return True

    def isPoint(self):
        return np.all(np.abs(self.interval[:, 0] - self.interval[:, 1]) < 1e-32)

    def startFinalStep(self):
        self.finalStep = True
        self.preFinalInterval = self.interval.copy()
        self.preFinalTransforms = self.transforms.copy()

    def getIntervalForCombining(self):
        return self.preFinalInterval if self.finalStep else self.interval

    def __repr__(self):
        return str(self.interval)


TrackedInterval = Box   # the reference's name


# ----------------------------------------------------------------------------
# linear-system box shrinking (:578-753)
# ----------------------------------------------------------------------------
def linear_terms(M):
    """getLinearTerms :578-605 -- coefficient of T_1(x_d) for every axis d (0 if extent 1)."""
    flat = M.ravel()
    out = []
    spot = 1
    for n in M.shape[::-1]:
        out.append(0 if n == 1 else flat[spot])
        spot *= n
    return out[::-1]


def bounding_interval_linear_system(Ms, errors, finalStep, macheps=MACHEPS):
    """BoundingIntervalLinearSystem :628-753.

    Returns (interval[dim,2], changed, should_stop, throwOut).
    """
    if finalStep:
        errors = np.zeros_like(errors)
    dim = Ms[0].ndim
    shrink_floor = 0.99                      # minZoomForChange
    basecase_floor = 0.4 ** dim              # minZoomForBaseCaseEnd
    A = np.array([linear_terms(M) for M in Ms])
    consts = np.array([M.ravel()[0] for M in Ms])
    total = np.array([np.sum(np.abs(M)) + e for M, e in zip(Ms, errors)])
    lin = np.sum(np.abs(A), axis=1)
    err = np.array([t - abs(c) - l for t, c, l in zip(total, consts, lin)])

    errors = errors.copy()
    for i in range(dim):                                                  # :669-678
        big = np.max(np.abs(A[i]))
        if big > 0:
            s = 2. ** int(np.floor(np.log2(abs(big))))
            A[i] /= s
            consts[i] /= s
            total[i] /= s
            lin[i] /= s
            err[i] /= s
            errors[i] /= s
    col_scale = np.ones(dim)
    for i in range(dim):                                                  # :680-687
        big = np.max(np.abs(A[:, i]))
        if big > 0:
            s = 2 ** (-np.floor(np.log2(abs(big))))
            col_scale[i] = s
            total += np.abs(A[:, i]) * (s - 1)
            A[:, i] *= s

    U, S, Vh = np.linalg.svd(A)                                           # :692-698
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        cond = S[-1] / S[0]
        well = bool(S[0] > 0 and cond > 1e-10)
        pad = max(cond, 2) * macheps
        Ainv = (1 / S * Vh.T) @ U.T
        center = -Ainv @ consts
    lo0, hi0 = kernels.linear_check(total, A, consts)                     # :700
    force_stop = finalStep and not well
    keep_lo = keep_hi = None
    for attempt in range(2):
        # NB the reference updates these arrays in place (:713-725).  When the
        # system is ill conditioned `lo`/`hi` alias the linear-check result, so
        # the second pass starts from the already scaled/padded/clipped bounds
        # and `keep_*` (aliases again) see that second round too.  Kept as is.
        lo, hi = lo0, hi0
        if well:
            width = np.sum(np.abs(Ainv * err), axis=1)
            lo = np.maximum(center - width, lo)
            hi = np.minimum(center + width, hi)
        lo *= col_scale
        hi *= col_scale
        lo -= pad
        hi += pad
        throw = bool(np.any(lo > hi) or np.any(lo > 1) or np.any(hi < -1))
        lo[lo < -1] = -1
        hi[hi < -1] = -1
        lo[lo > 1] = 1
        hi[hi > 1] = 1
        ratio = np.prod(hi - lo) / 2 ** dim
        if throw:
            changed = True
        elif attempt == 0:
            changed = ratio < shrink_floor
        else:
            changed = ratio < basecase_floor
        if attempt == 0 and changed:
            return np.vstack([lo, hi]).T, changed, force_stop, throw
        if attempt == 0:
            keep_lo, keep_hi = lo, hi
            err = errors
        elif changed:
            return np.vstack([keep_lo, keep_hi]).T, False, force_stop, False
        else:
            return np.vstack([keep_lo, keep_hi]).T, False, well or force_stop, False


# ----------------------------------------------------------------------------
# interval restriction of whole systems (:812-894)
# ----------------------------------------------------------------------------
def transformation_error(M, axis):
    """getTransformationError :812-833 -- n_axis * eps * sum|M|."""
    return M.shape[axis] * MACHEPS * np.sum(np.abs(M))


def restrict_tensor(M, alphas, betas, error, exact):
    """transformCheb :835-862 -- every axis in turn, error charged before each axis."""
    for axis, alpha, beta in zip(range(M.ndim), alphas, betas):
        error += transformation_error(M, axis)
        M = kernels.restrict_axis(M, axis, alpha, beta, exact)
    return M, error


def restrict_system(Ms, alphas, betas, errors, exact):
    """transformChebToInterval :864-894."""
    out, errs = [], []
    for M, e in zip(Ms, errors):
        M2, e2 = restrict_tensor(M, alphas, betas, e, exact)
        out.append(M2)
        errs.append(e2)
    return out, np.array(errs)


def zoom_once(Ms, errors, box, exact):
    """zoomInOnIntervalIter :896-956.

    Returns (Ms, errors, box, changed, should_stop).
    """
    sub, changed, should_stop, throw = bounding_interval_linear_system(Ms, errors, box.finalStep)
    for d in range(len(Ms)):                          # frozen (degenerate) axes :932-935
        if box.interval[d, 0] == box.interval[d, 1]:
            sub[d, 0] = -1.
            sub[d, 1] = 1.
    if throw and not box.canThrowOut():               # :937-940
        throw = False
        should_stop = True
        changed = True
    if throw:
        box.empty = True
        return Ms, errors, box, True, True
    if not changed:
        return Ms, errors, box, changed, should_stop
    box.addTransform(sub)
    Ms, errors = restrict_system(Ms, *box.getLastTransform(), errors, exact)
    if box.finalStep and box.isPoint():               # :952-954
        should_stop = True
        changed = False
    return Ms, errors, box, changed, should_stop


# ----------------------------------------------------------------------------
# subdivision (:981-1129)
# ----------------------------------------------------------------------------
def inverse_order(order):
    """getInverseOrder :981-1011.

    `order` is the sequence of axes one tensor was halved along.  The 2^k
    pieces come out lower-half-first per axis in that sequence; the returned
    permutation lists them as if the axes had been halved in ascending order.
    """
    rank = np.zeros_like(order)
    rank[np.argsort(order)] = np.arange(len(rank))
    weight = 2 ** (len(rank) - 1 - rank)
    pos = np.array([bits @ weight for bits in itertools.product([0, 1], repeat=len(weight))])
    inv = np.zeros_like(pos)
    inv[pos] = np.arange(len(pos))
    return tuple(inv)


def subdivision_axes(Ms, box, level):
    """getSubdivisionDims :1013-1049 -- row i = axes (in order) to halve tensor i along."""
    dim = len(Ms)
    cand = np.arange(dim)
    for i in range(dim):
        if np.isclose(box.interval[i, 0], box.interval[i, 1]):
            if len(cand) != 1:
                cand = np.delete(cand, np.argwhere(cand == i))
    if level <= 5:
        widths = box.dimSize()
        widest = np.max([widths[i] for i in cand])
        cand = np.extract(widths[cand] > widest / 5, cand)
        if len(cand) > 1:
            shapes = np.array([np.array(M.shape) for M in Ms])
            per_axis = np.sum(shapes, axis=0)
            total = np.sum(per_axis)
            for i in cand.copy():
                if len(cand) > 1 and per_axis[i] < np.floor(total / (dim + 1)):
                    cand = np.delete(cand, np.argwhere(cand == i))
    return np.vstack([cand[np.argsort(np.array(M.shape)[cand])[::-1]] for M in Ms])


def subdivide(Ms, errors, box, exact, level):
    """getSubdivisionIntervals :1051-1129 -> (child tensors, child errors, child boxes)."""
    axes = subdivision_axes(Ms, box, level)
    axis_set = set(axes.flatten())
    if len(axis_set) != axes.shape[1]:
        raise ValueError("Subdivision Dimensions are invalid! Each Polynomial must subdivide in the same dimensions!")
    per_poly_M, per_poly_E = [], []
    for M, error, order in zip(Ms, errors, axes):
        pieces, perrs = [M], [error]
        for axis in order:
            mid = box.nextTransformPoints[axis]
            alpha, beta = (mid + 1) / 2, (mid - 1) / 2
            nxt, nxt_e = [], []
            for T, E in zip(pieces, perrs):
                lower = kernels.restrict_axis(T, axis, alpha, beta, exact)
                upper = kernels.restrict_axis(T, axis, -beta, alpha, exact)
                e1 = transformation_error(T, axis)
                nxt += [lower, upper]
                nxt_e += [e1 + E, e1 + E]
            pieces, perrs = nxt, nxt_e
        if M.ndim == 1:
            per_poly_M.append(pieces)
            per_poly_E.append(perrs)
        else:
            inv = inverse_order(order)
            per_poly_M.append([pieces[i] for i in inv])
            per_poly_E.append([perrs[i] for i in inv])
    nchild = len(per_poly_M[0])
    child_Ms = [[per_poly_M[p][c] for p in range(len(per_poly_M))] for c in range(nchild)]
    child_Es = [[per_poly_E[p][c] for p in range(len(per_poly_E))] for c in range(nchild)]
    boxes = [box]
    for axis in axis_set:
        mid = box.nextTransformPoints[axis]
        sub = np.ones_like(box.interval)
        sub[:, 0] = -1.
        grown = []
        for old in boxes:
            lo_half = old.copy()
            hi_half = old.copy()
            sub[axis] = [-1., mid]
            lo_half.addTransform(sub)
            sub[axis] = [mid, 1.]
            hi_half.addTransform(sub)
            lo_half.nextTransformPoints[axis] = 0
            hi_half.nextTransformPoints[axis] = 0
            grown += [lo_half, hi_half]
        boxes = grown
    return child_Ms, child_Es, boxes


def trim_system(Ms, errors, relApproxTol=1e-3, absApproxTol=0):
    """trimMs :1131-1171 -- drop negligible trailing hyperplanes (in place on the lists)."""
    dim = Ms[0].ndim
    for p in range(len(Ms)):
        budget = absApproxTol + errors[p] * relApproxTol
        sel = [slice(None)] * dim
        for axis in range(dim):
            sel[axis] = -1
            last = np.sum(np.abs(Ms[p][tuple(sel)]))
            while last < budget and Ms[p].shape[axis] > 3:
                sel[axis] = slice(None, -1)
                Ms[p] = Ms[p][tuple(sel)]
                budget -= last
                errors[p] += last
                sel[axis] = -1
                last = np.sum(np.abs(Ms[p][tuple(sel)]))
            sel[axis] = slice(None)


# ---- the alternative trim policies of the reference's experiment file ------------------------------------------------
# /root/reference/tests/TrimMs optimization/ChebyshevSubdivisionSolverWithOptimizedTrimMs.py (cited below as T:line).  That
# file is an older copy of the solver in which the experimenter swaps the function called at the trimMs call site by hand;
# the three functions themselves are what is restated here (pinned bit for bit: tests/golden/trim_policies.npz), and
# `trim_policy` plugs them into the CURRENT solver's call site (:1232).
def trim_system_optimized1(Ms, errors, relApproxTol=1e-3, absApproxTol=0):
    """trimMsOptimized1 T:1167-1197 -- one hyperplane per axis per sweep, while anything can be trimmed (extent > 2)."""
    dim = Ms[0].ndim
    for p in range(len(Ms)):
        budget = absApproxTol + errors[p] * relApproxTol
        again = True
        while again:
            again = False
            swept = 0
            for axis in range(dim):
                last = np.sum(np.abs(np.take(Ms[p], indices=-1, axis=axis)))
                if Ms[p].shape[axis] > 2 and last < budget:
                    again = True
                    Ms[p] = np.delete(Ms[p], -1, axis=axis)
                    budget -= last
                    swept += last
            errors[p] += swept


def trim_system_optimized2(Ms, errors, relApproxTol=1e-3, absApproxTol=0):
    """trimMsOptimized2 T:1199-1235 -- always the axis whose last hyperplane is the smallest."""
    dim = Ms[0].ndim
    for p in range(len(Ms)):
        budget = absApproxTol + errors[p] * relApproxTol
        while True:
            best, best_axis = float("inf"), None
            for axis in range(dim):
                last = np.sum(np.abs(np.take(Ms[p], indices=-1, axis=axis)))
                if last < best and Ms[p].shape[axis] > 2:
                    best, best_axis = last, axis
            if best_axis is None or best >= budget:
                break
            Ms[p] = np.delete(Ms[p], -1, axis=best_axis)
            budget -= best
            errors[p] += best


def trim_system_optimized3(Ms, errors, relApproxTol=1e-3, absApproxTol=0):
    """trimMsOptimized3 T:1237-1271 -- axes by descending extent, each trimmed as far as it goes.  (The running total is
    added to the error after EVERY axis, T:1271 sits inside the loop: the error grows by more than what was removed.)"""
    for p in range(len(Ms)):
        budget = absApproxTol + errors[p] * relApproxTol
        total = 0
        for axis in np.argsort(Ms[p].shape)[::-1]:
            while True:
                last = np.sum(np.abs(np.take(Ms[p], indices=-1, axis=axis)))
                if last < budget and Ms[p].shape[axis] > 2:
                    Ms[p] = np.delete(Ms[p], -1, axis=axis)
                    budget -= last
                    total += last
                else:
                    break
            errors[p] += total


TRIM_POLICIES = {0: trim_system, 1: trim_system_optimized1, 2: trim_system_optimized2, 3: trim_system_optimized3}


def is_exterior(entry_box, box):
    """isExteriorInterval :1173-1175."""
    return np.any(box.getIntervalForCombining() == entry_box.getIntervalForCombining())


# ----------------------------------------------------------------------------
# the recursion (:1177-1385) and the entry point (:1387-1464)
# ----------------------------------------------------------------------------
class Options:
    """SolverOptions :12-43 plus the ad-hoc useFinalStep (:1430)."""

    def __init__(self):
        self.verbose = False
        self.exact = False
        self.constant_check = True
        self.low_dim_quadratic_check = True
        self.all_dim_quadratic_check = False
        self.maxZoomCount = 25
        self.level = 0
        self.useFinalStep = True
        self.trim_policy = 0           # 0: trimMs :1131; 1-3: trimMsOptimized1/2/3 of the reference's experiment file


class Trace:
    """Counters the benchmark and the parity tests read (not in the reference)."""

    def __init__(self):
        self.boxes = 0            # invocations of solve_box == reference solvePolyRecursive
        self.zooms = 0
        self.subdivisions = 0
        self.merges = 0
        self.discard_const = 0
        self.discard_quad = 0
        self.discard_zoom = 0
        self.entry_coeffs = 0     # sum over boxes of sum_p prod(shape) at entry (SURVEY 8d bytes)
        self.restrict_flops = 0   # sum (n_axis+1)*N_c over restrictions   (SURVEY 8d flops)
        self.per_level = {}


def solve_box(Ms, box, errors, opts, trace):
    """solvePolyRecursive :1177-1385 -> (interior boxes, exterior boxes)."""
    trace.boxes += 1
    if box.isPoint():
        return [], [box]
    level = opts.level + 1
    trace.per_level[level] = trace.per_level.get(level, 0) + 1
    trace.entry_coeffs += sum(int(M.size) for M in Ms)

    if opts.constant_check:                                              # :1213-1217
        consts = np.array([M.ravel()[0] for M in Ms])
        rest = np.array([np.sum(np.abs(M)) - abs(c) + e for M, e, c in zip(Ms, errors, consts)])
        if np.any(np.abs(consts) > rest):
            trace.discard_const += 1
            return [], []
    if (opts.low_dim_quadratic_check and Ms[0].ndim <= 3) or opts.all_dim_quadratic_check:
        for i in range(len(Ms)):                                         # :1221-1224
            if quadratic_check(Ms[i], errors[i]):
                trace.discard_quad += 1
                return [], []

    Ms = list(Ms)
    entry_Ms = list(Ms)
    box = box.copy()
    errors = errors.copy()
    TRIM_POLICIES[getattr(opts, "trim_policy", 0)](Ms, errors)

    changed = True
    zoom_count = 0
    entry_box = box.copy()
    last = box.dimSize()
    should_stop = False
    while changed and zoom_count <= opts.maxZoomCount:                   # :1243-1252
        Ms, errors, box, changed, should_stop = zoom_once(Ms, errors, box, opts.exact)
        trace.zooms += 1
        if box.empty:
            trace.discard_zoom += 1
            return [], []
        now = box.dimSize()
        if np.all(now >= last / 2):
            zoom_count += 1
        last = now

    sub_opts = _child_options(opts, level)
    if should_stop:
        if box.finalStep or not opts.useFinalStep:                       # :1256-1262
            return ([], [box]) if is_exterior(entry_box, box) else ([box], [])
        box.startFinalStep()                                             # :1264-1265
        return solve_box(Ms, box, errors, sub_opts, trace)
    if box.finalStep:                                                    # :1266-1298
        box.canThrowOutFinalStep = True
        trace.subdivisions += 1
        kids_M, kids_E, kids_B = subdivide(Ms, errors, box, opts.exact, level)
        found = []
        for m, e, b in zip(kids_M, kids_E, kids_B):
            a, c = solve_box(m, b, np.array(e), sub_opts, trace)
            found += a + c
        if len(found) == 0:
            box.possibleExtraRoot = True
        else:
            seen = set()
            distinct = []
            for r in found:
                key = tuple(r.interval[:, 0])
                if key in seen:
                    continue
                seen.add(key)
                distinct.append(r)
            for r in distinct:
                if len(r.possibleDuplicateRoots) > 0:
                    box.possibleDuplicateRoots += r.possibleDuplicateRoots
                else:
                    box.possibleDuplicateRoots.append(r.getFinalPoint())
        return ([], [box]) if is_exterior(entry_box, box) else ([box], [])

    if level == 15:                                                      # :1302-1309
        warnings.warn("High subdivision depth!\nSubdivision on the search interval has now reached"
                      " at least depth 15. Runtime may be prolonged.")
    elif level == 25:
        warnings.warn("Extreme subdivision depth!\nSubdivision on the search interval has now reached"
                      " at least depth 25, which is unusual. The solver may not finish running."
                      "Ensure the input functions meet the requirements of being continuous, smooth,"
                      "and having only finitely many simple roots on the search interval.")
    trace.subdivisions += 1
    interior, exterior = [], []
    kids_M, kids_E, kids_B = subdivide(Ms, errors, box, opts.exact, level)
    for m, e, b in zip(kids_M, kids_E, kids_B):
        a, c = solve_box(m, b, np.array(e), sub_opts, trace)
        interior += a
        exterior += c

    # merge exterior boxes that touch / overlap, re-solve the merged ones (:1318-1385)
    for b in exterior:
        b.reRun = False
    i, j = 0, 1
    while i < len(exterior):
        while j < len(exterior):
            if exterior[i].overlapsWith(exterior[j]):
                trace.merges += 1
                merged = entry_box.copy()
                if merged.finalStep:
                    merged.interval = merged.preFinalInterval.copy()
                    merged.transforms = merged.preFinalTransforms.copy()
                Ia, Ib = exterior[i].getIntervalForCombining(), exterior[j].getIntervalForCombining()
                new_lo = np.min([Ia[:, 0], Ib[:, 0]], axis=0)
                new_hi = np.max([Ia[:, 1], Ib[:, 1]], axis=0)
                Fa, Fb = exterior[i].getFinalInterval(), exterior[j].getFinalInterval()
                new_lo_f = np.min([Fa[:, 0], Fb[:, 0]], axis=0)
                new_hi_f = np.max([Fa[:, 1], Fb[:, 1]], axis=0)
                old_lo = entry_box.interval[:, 0]
                old_hi = entry_box.interval[:, 1]
                old_lo_f, old_hi_f = entry_box.getFinalInterval().T
                flat = old_hi_f == old_lo_f
                old_hi_f[flat] = old_hi_f[flat] + 1
                sub = ((2 * np.array([new_lo_f, new_hi_f]) - old_lo_f - old_hi_f) / (old_hi_f - old_lo_f)).T
                sub[flat, 0] = -1
                sub[flat, 1] = 1
                sub[:, 0][old_lo == new_lo] = -1
                sub[:, 1][old_hi == new_hi] = 1
                merged.addTransform(sub)
                merged.interval = np.array([new_lo, new_hi]).T
                merged.reRun = True
                del exterior[j]
                del exterior[i]
                exterior.append(merged)
                j = i + 1
            else:
                j += 1
        i += 1
        j = i + 1
    still_exterior = []
    for b in exterior:
        if b.reRun:
            if np.all(b.interval == entry_box.interval):
                still_exterior.append(b)
            else:
                m2, e2 = restrict_system(entry_Ms, *b.getLastTransform(), errors, opts.exact)
                a, c = solve_box(m2, b, e2, sub_opts, trace)
                interior += a
                still_exterior += c
        elif is_exterior(entry_box, b):
            still_exterior.append(b)
        else:
            interior.append(b)
    return interior, still_exterior


def _child_options(opts, level):
    import copy
    o = copy.copy(opts)
    o.level = level
    return o


def solve_chebyshev_subdivision(Ms, errors, verbose=False, returnBoundingBoxes=False, exact=False,
                                constant_check=True, low_dim_quadratic_check=True,
                                all_dim_quadratic_check=False, trace=None, trim_policy=0):
    """solveChebyshevSubdivision :1387-1464.  (`trim_policy`: see TRIM_POLICIES; not a parameter of the reference.)"""
    if np.any([M.ndim != len(Ms) for M in Ms]):
        raise ValueError("Solver Takes in N polynomials of dimension N!")
    if len(Ms) != len(errors):
        raise ValueError("Ms and errors must be same length!")
    top = Box(np.array([[-1., 1.]] * Ms[0].ndim))
    opts = Options()
    opts.verbose = verbose
    opts.exact = exact
    opts.constant_check = constant_check
    opts.low_dim_quadratic_check = low_dim_quadratic_check
    opts.all_dim_quadratic_check = all_dim_quadratic_check
    opts.trim_policy = trim_policy
    if trace is None:
        trace = Trace()
    inner, outer = solve_box(Ms, top, errors, opts, trace)
    boxes = inner + outer
    roots = []
    dup = extra = False
    for b in boxes:
        b.getFinalInterval()
        if b.possibleExtraRoot:
            extra = True
        if len(b.possibleDuplicateRoots) > 0:
            roots += b.possibleDuplicateRoots
            dup = True
        else:
            roots.append(b.getFinalPoint())
    if extra:
        warnings.warn("Might Have Extra Roots! See Bounding Boxes for details!")
    if dup:
        warnings.warn("Might Have Duplicate Roots! See Bounding Boxes for details!")
    roots = np.array(roots)
    if returnBoundingBoxes:
        return roots, boxes
    return roots


solveChebyshevSubdivision = solve_chebyshev_subdivision   # the reference's name
