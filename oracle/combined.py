"""ORACLE -- test infrastructure only (see oracle/README.md).

CPU restatement of /root/reference/yroots/Combined_Solver.py:10-200 (`solve`):
approximate every function, run the subdivision solver, re-solve boxes that are
still larger than minBoundingIntervalSize, map everything back to [a, b].
"""
import functools
import itertools

import numpy as np

from . import approximator, subdivision
from .polynomial import MultiCheb, MultiPower

RESPLIT = 0.51234912839471234       # :146


def solve(funcs, a=-1, b=1, verbose=False, returnBoundingBoxes=False, exact=False,
          minBoundingIntervalSize=1e-5, trace=None):
    if type(funcs) != list and type(funcs) != np.ndarray:
        funcs = [funcs]
    for i, f in enumerate(funcs):
        if not hasattr(f, '__call__'):
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
    if trace is None:
        trace = subdivision.Trace()
    unit = np.allclose(a, -np.ones_like(a)) and np.allclose(b, np.ones_like(b))
    polys = [None] * dim
    errs = np.zeros(dim)
    for i, f in enumerate(funcs):                                     # :117-129
        if unit and isinstance(f, MultiPower):
            polys[i], errs[i] = f.to_cheb(), 2.0 ** -52
        elif unit and isinstance(f, MultiCheb):
            polys[i], errs[i] = f.coeff, 2.0 ** -52
        else:
            polys[i], errs[i] = approximator.cheb_approximate(f, a, b)
    kw = dict(verbose=verbose, returnBoundingBoxes=True, exact=exact,
              minBoundingIntervalSize=minBoundingIntervalSize, trace=trace)
    roots, boxes = subdivision.solve_chebyshev_subdivision(
        polys, errs, verbose, True, exact, constant_check=True,
        low_dim_quadratic_check=True, all_dim_quadratic_check=False, trace=trace)

    splitting_allowed = np.all(b - a > minBoundingIntervalSize)
    if len(boxes) == 1 and np.all(boxes[0].finalDimSize() == 2) and splitting_allowed:   # :138-162
        roots, out_boxes = [], []
        for upper in itertools.product([False, True], repeat=len(a)):
            mid = (a + b) * RESPLIT
            r, bx = solve(funcs, a=np.where(upper, mid, a), b=np.where(upper, b, mid), **kw)
            if len(r) != 0:
                out_boxes.append(bx)
                roots.append(r)
        if len(roots) > 0:
            roots = np.vstack(roots)
            out_boxes = np.vstack(out_boxes)
        return (roots, out_boxes) if returnBoundingBoxes else roots

    out_roots, out_boxes = [], []
    for box in boxes:                                                 # :170-190
        lo, hi = approximator.affine_to_box(box.finalInterval.T, a, b)
        limit = minBoundingIntervalSize * functools.reduce(np.maximum, [np.abs(a), np.abs(b), 1])
        if np.all(hi - lo > limit):
            r, bx = solve(funcs, a=lo, b=hi, **kw)
            if len(r) > 0:
                out_roots.append(r)
                out_boxes.append(bx)
        else:
            out_boxes.append([approximator.affine_to_box(box.finalInterval.T, a, b).T])
            if len(box.possibleDuplicateRoots) > 0:
                out_roots.append(approximator.affine_to_box(np.array(box.possibleDuplicateRoots), a, b))
            else:
                out_roots.append(approximator.affine_to_box(box.getFinalPoint(), a, b))
    if len(out_boxes) != 0:
        out_boxes = np.vstack(out_boxes)
    if len(out_roots) != 0:
        out_roots = np.vstack(out_roots)
    return (out_roots, out_boxes) if returnBoundingBoxes else out_roots
