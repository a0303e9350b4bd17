"""Drop-in for yroots.Combined_Solver.solve (/root/reference/yroots/Combined_Solver.py:10-200).

Host orchestration only: validate -> approximate every function (ChebyshevApproximator, device DCT)
-> subdivision solve (device frontier) -> re-solve boxes still wider than minBoundingIntervalSize
-> map roots and boxes back to [a, b].  Signature, return types (including the Python list `[]`
when nothing is found) and ValueError strings are the reference's.
"""
import functools
import itertools

import numpy as np

from . import ChebyshevApproximator, ChebyshevSubdivisionSolver
from .polynomial import MultiCheb, MultiPower
from .runtime import get_engine

_RESPLIT = 0.51234912839471234      # :146


class _Region:
    """One search box of the recursion of /root/reference/yroots/Combined_Solver.py:138-190.  `parts` lists, in the
    reference's reporting order, what the region contributes: finished (roots, boxes) pairs and sub-regions."""
    __slots__ = ("a", "b", "parts")

    def __init__(self, a, b):
        self.a, self.b, self.parts = a, b, []

    def collect(self, roots, boxes):
        for part in self.parts:
            if isinstance(part, _Region):
                part.collect(roots, boxes)
            else:
                roots.append(part[0])
                boxes.append(part[1])


def _validated(funcs, a, b):
    if type(funcs) != list and type(funcs) != np.ndarray:
        funcs = [funcs]
    for i in range(len(funcs)):
        if not hasattr(funcs[i], '__call__'):
            raise ValueError(f"Invalid input: input function {i} is not callable")
    dim = len(funcs)
    if type(a) == list:
        a = np.array(a)
    if type(b) == list:
        b = np.array(b)
    if type(a) != np.ndarray:
        a = np.full(dim, a)
    if type(b) != np.ndarray:
        b = np.full(dim, b)
    if len(a) != len(b):
        raise ValueError(f"Invalid input: {len(a)} lower bounds were given but {len(b)} upper bounds were given")
    if (b < a).any():
        raise ValueError("Invalid input: at least one lower bound is greater than the corresponding upper bound.")
    return funcs, a, b


def _approximate_level(funcs, level, verbose):
    """Chebyshev tensors + error bounds of every function on every region of a recursion level (:108-129).  Polynomial
    objects on the unit box are used as they are; everything else goes through ONE batched approximation whose degree
    searches advance in lock step on the device (ChebyshevApproximator.chebApproximate_batch)."""
    dim = len(funcs)
    out, jobs, where = [], [], []
    for r, reg in enumerate(level):
        unit_box = np.allclose(reg.a, -np.ones_like(reg.a)) and np.allclose(reg.b, np.ones_like(reg.b))
        polys, errs = [None] * dim, np.zeros(dim)
        for i in range(dim):
            if unit_box and isinstance(funcs[i], MultiPower):
                polys[i], errs[i] = funcs[i].to_cheb(engine=get_engine()), 2.0 ** -52
            elif unit_box and isinstance(funcs[i], MultiCheb):
                polys[i], errs[i] = funcs[i].coeff, 2.0 ** -52
            else:
                jobs.append((funcs[i], reg.a, reg.b))
                where.append((r, i))
        out.append((polys, errs))
    if jobs:
        for (r, i), (coeff, err) in zip(where, ChebyshevApproximator.chebApproximate_batch(jobs)):
            out[r][0][i], out[r][1][i] = coeff, err
    if verbose:
        for reg, (polys, errs) in zip(level, out):
            print("Approximation shapes:", " ".join(f"{i}: {polys[i].shape}" for i in range(dim)))
            print(f"Searching on interval {[[reg.a[i], reg.b[i]] for i in range(dim)]}")
    return out


def solve(funcs, a=-1, b=1, verbose=False, returnBoundingBoxes=False, exact=False, minBoundingIntervalSize=1e-5):
    """Roots of the system `funcs` on the box [a, b] (and optionally a bounding box per root).

    The reference recurses: a result box still wider than minBoundingIntervalSize is re-approximated and re-solved on its
    own (:170-183), and a solve that only returns the whole search box is split into 2^dim pieces (:138-162) -- one
    approximation and one device solve per box, depth first.  Here the recursion runs LEVEL BY LEVEL: every region of a
    level is approximated, all of them are solved in ONE device batch (solve_batch), and their oversized boxes become
    the regions of the next level; a solve makes O(depth) device calls instead of O(roots).  Regions keep their place in
    a tree, so roots and boxes come back in the reference's (depth-first) order.
    """
    funcs, a, b = _validated(funcs, a, b)
    dim = len(funcs)
    top = _Region(a, b)
    level = [top]
    while level:
        systems, errors = [], []
        if verbose:
            for reg in level:
                if reg is not top:
                    print("Re-solving on:", reg.a, reg.b)
        for polys, errs in _approximate_level(funcs, level, verbose):
            ChebyshevSubdivisionSolver._check_system(polys, errs)
            systems.append(polys)
            errors.append(errs)
        if verbose:
            print(f"Finding roots on {len(level)} region(s)...")
        found, boxes_all = ChebyshevSubdivisionSolver.solve_batch(systems, errors, True, exact, True, True, False)
        nxt = []
        for reg, boxes in zip(level, boxes_all):
            lo0, hi0 = reg.a, reg.b
            if len(boxes) == 1 and np.all(boxes[0].finalDimSize() == 2) and np.all(hi0 - lo0 > minBoundingIntervalSize):
                # the whole region came back as one box: cut it (almost) in half along every axis and solve the pieces
                mid = (lo0 + hi0) * _RESPLIT
                for upper in itertools.product([False, True], repeat=dim):
                    sub = _Region(np.where(upper, mid, lo0), np.where(upper, hi0, mid))
                    reg.parts.append(sub)
                    nxt.append(sub)
                continue
            limit = minBoundingIntervalSize * functools.reduce(np.maximum, [np.abs(lo0), np.abs(hi0), 1])
            for box in boxes:
                lo, hi = ChebyshevApproximator.transform(box.finalInterval.T, lo0, hi0)
                if np.all(hi - lo > limit):
                    sub = _Region(lo, hi)                              # still too wide: its own approximation and solve
                    reg.parts.append(sub)
                    nxt.append(sub)
                    continue
                pts = np.array(box.possibleDuplicateRoots) if len(box.possibleDuplicateRoots) > 0 else box.getFinalPoint()
                reg.parts.append((np.atleast_2d(ChebyshevApproximator.transform(pts, lo0, hi0)),
                                  [np.stack([lo, hi], axis=1)]))
        if verbose:
            print(f"Found {sum(len(r) for r in found)} roots; {len(nxt)} region(s) to re-solve")
        level = nxt
    roots, boxes = [], []
    top.collect(roots, boxes)
    if len(boxes) != 0:
        boxes = np.vstack(boxes)
    if len(roots) != 0:
        roots = np.vstack(roots)
    return (roots, boxes) if returnBoundingBoxes else roots
