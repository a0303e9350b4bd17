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

_RESPLIT = 0.51234912839471234      # :146


def solve(funcs, a=-1, b=1, verbose=False, returnBoundingBoxes=False, exact=False, minBoundingIntervalSize=1e-5):
    """Roots of the system `funcs` on the box [a, b] (and optionally a bounding box per root)."""
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
    unit_box = np.allclose(a, -np.ones_like(a)) and np.allclose(b, np.ones_like(b))
    polys, errs = [None] * dim, np.zeros(dim)
    if verbose:
        print("Approximation shapes:", end=" ")
    for i in range(dim):                                                  # :117-129
        if unit_box and isinstance(funcs[i], MultiPower):
            polys[i], errs[i] = funcs[i].to_cheb(), 2.0 ** -52
        elif unit_box and isinstance(funcs[i], MultiCheb):
            polys[i], errs[i] = funcs[i].coeff, 2.0 ** -52
        else:
            polys[i], errs[i] = ChebyshevApproximator.chebApproximate(funcs[i], a, b)
        if verbose:
            print(f"{i}: {polys[i].shape}", end=" " if i != dim - 1 else '\n')
    if verbose:
        print(f"Searching on interval {[[a[i], b[i]] for i in range(dim)]}")

    roots, boxes = ChebyshevSubdivisionSolver.solveChebyshevSubdivision(
        polys, errs, verbose, True, exact, constant_check=True, low_dim_quadratic_check=True,
        all_dim_quadratic_check=False)
    again = dict(verbose=verbose, returnBoundingBoxes=True, exact=exact, minBoundingIntervalSize=minBoundingIntervalSize)

    may_split = np.all(b - a > minBoundingIntervalSize)
    if len(boxes) == 1 and np.all(boxes[0].finalDimSize() == 2) and may_split:       # :138-162
        all_roots, all_boxes = [], []
        for upper in itertools.product([False, True], repeat=len(a)):
            mid = (a + b) * _RESPLIT
            lo, hi = np.where(upper, mid, a), np.where(upper, b, mid)
            if verbose:
                print("Re-solving on:", lo, hi)
            r, bx = solve(funcs, a=lo, b=hi, **again)
            if len(r) != 0:
                all_boxes.append(bx)
                all_roots.append(r)
        if len(all_roots) > 0:
            all_roots, all_boxes = np.vstack(all_roots), np.vstack(all_boxes)
        return (all_roots, all_boxes) if returnBoundingBoxes else all_roots

    final_roots, final_boxes = [], []
    for box in boxes:                                                     # :170-190
        lo, hi = ChebyshevApproximator.transform(box.finalInterval.T, a, b)
        limit = minBoundingIntervalSize * functools.reduce(np.maximum, [np.abs(a), np.abs(b), 1])
        if np.all(hi - lo > limit):
            if verbose:
                print("Re-solving on:", lo, hi)
            r, bx = solve(funcs, a=lo, b=hi, **again)
            if len(r) > 0:
                final_roots.append(r)
                final_boxes.append(bx)
        else:
            final_boxes.append([ChebyshevApproximator.transform(box.finalInterval.T, a, b).T])
            if len(box.possibleDuplicateRoots) > 0:
                final_roots.append(ChebyshevApproximator.transform(np.array(box.possibleDuplicateRoots), a, b))
            else:
                final_roots.append(ChebyshevApproximator.transform(box.getFinalPoint(), a, b))
    if len(final_boxes) != 0:
        final_boxes = np.vstack(final_boxes)
    if len(final_roots) != 0:
        final_roots = np.vstack(final_roots)
    return (final_roots, final_boxes) if returnBoundingBoxes else final_roots
