"""ORACLE -- test infrastructure only (see oracle/README.md).

CPU restatement of /root/reference/yroots/ChebyshevApproximator.py: sampling on
Chebyshev-Lobatto grids, DCT-I -> coefficients, and the coefficient-decay
degree search.  The DCT itself is scipy.fftpack.dctn(type=1), the same
third-party routine the reference calls (:78; pinned scipy==1.10.1 in
requirements.txt:2, 1.18.1 in this image).
"""
import itertools
import warnings

import numpy as np
from scipy.fftpack import dctn

from .polynomial import MultiCheb, MultiPower

MACHEPS = 2.0 ** -52


def affine_to_box(x, a, b):
    """transform :8-26 -- [-1,1] -> [a,b]."""
    return ((b - a) * x + (b + a)) / 2


def sample_and_dct(f, degs, a, b, retSupNorm=False):
    """interval_approximate_nd :28-89."""
    dim = len(degs)
    wanted = degs.copy()
    degs[degs == 0] = 1                               # NB mutates the caller's array, as :57 does
    axes = [affine_to_box(np.cos(np.arange(n + 1) * np.pi / n), lo, hi)
            for n, lo, hi in zip(degs, a, b)]
    grid = np.meshgrid(*axes, indexing='ij')
    pts = np.column_stack(tuple(g.flatten() for g in grid))
    if isinstance(f, (MultiCheb, MultiPower)) or getattr(f, "_takes_point_array", False):
        values = f(pts).reshape(*(degs + 1))
    else:
        values = np.array([f(*p) for p in pts]).reshape(*(degs + 1))
    sup = np.max(np.abs(values)) if retSupNorm else None
    coeffs = dctn(values / np.prod(degs), type=1, overwrite_x=True)
    for d in range(dim):                              # halve the two boundary hyperplanes per axis
        first = tuple(slice(None) if i != d else 0 for i in range(dim))
        last = tuple(slice(None) if i != d else degs[i] for i in range(dim))
        coeffs[first] /= 2
        coeffs[last] /= 2
    out = coeffs[tuple(slice(0, n + 1) for n in wanted)]
    return (out, sup) if retSupNorm else out


def started_converging(chunk, tol):
    """:91-106 -- the last five axis-mean |coeff| are below tol."""
    return np.all(chunk[-5:] < tol)


def has_converged(coeff, coeff2, tol):
    """:108-129 -- degree n and degree 2n+1 tensors agree to tol (missing entries count as 0)."""
    diff = coeff2.copy()
    diff[tuple(slice(0, n) for n in coeff.shape)] -= coeff
    return np.max(np.abs(diff)) < tol


def final_degree(chunk, tol, macheps=MACHEPS):
    """getFinalDegree :131-172 -> (degree, epsVal, rho)."""
    settled = int(3 * (len(chunk) - 1) / 4)
    eps = 2 * max(macheps, np.max(chunk[settled:]))
    big = np.where(chunk > eps)[0]
    degree = 1 if len(big) == 0 else max(1, big[-1])
    if np.all(chunk[1:] < tol):
        degree = 0
    peak = np.argmax(chunk)
    if eps == 0:
        eps = chunk[peak] * 1e-24
    rho = (chunk[peak] / eps) ** (1 / ((degree - peak) + 1))
    return degree, eps, rho


_PROBE_A = [-0.7996847717584993, 0.18546110255464776, -0.13975937255055182, 0., 1., -1.]
_PROBE_B = [-0.17223860129797386, 0.10828286380141305, -0.5333148248321931, 0.46471703497219596]


def constant_along(f, a, b, axis, relTol, absTol=0):
    """checkConstantInDimension :174-226."""
    poly = isinstance(f, (MultiPower, MultiCheb)) or getattr(f, "_takes_point_array", False)
    call = (lambda p: f(p)) if poly else (lambda p: f(*p))
    dim = len(a)
    for base, probes in (
            (np.array([0.8984743990614998 ** (k + 1) for k in range(dim)]), _PROBE_A),
            (np.array([(-0.2598647169391334 * (k + 1) / dim) ** 2 for k in range(dim)]), _PROBE_B)):
        x = affine_to_box(base, a, b)
        ref = call(x)
        if np.isclose(ref, 0, rtol=relTol, atol=absTol):
            return False
        for v in affine_to_box(np.array(probes), a[axis], b[axis]):
            x[axis] = v
            if not np.allclose(ref, call(x), rtol=relTol, atol=absTol):
                return False
    return True


def chebyshev_degrees(f, a, b, relTol, absTol=0):
    """getChebyshevDegrees :228-311 -> (degrees, epsilons, rhos)."""
    dim = len(a)
    found = [np.inf] * dim
    eps_list, rho_list = [], []
    for axis in range(dim):
        if constant_along(f, a, b, axis, relTol, absTol):
            found[axis] = 0
    for axis in range(dim):
        if found[axis] == 0:
            eps_list.append(0)
            rho_list.append(np.inf)
            continue
        degs = np.array([5] * dim if dim <= 5 else [2] * dim)
        for i in range(dim):
            if found[i] < degs[i]:
                degs[i] = found[i]
        guess = 8
        others = tuple(i for i in range(dim) if i != axis)
        while True:
            if guess > 1e5:
                warnings.warn(f"Approximation bound exceeded!\n\nApproximation degree in dimension {axis} "
                              "has exceeded 1e5, so the process may not finish.\n\nConsider interrupting "
                              "and restarting the process after ensuring that the function(s) inputted are "
                              "continuous and smooth on the approximation interval.\n\n")
            degs[axis] = guess
            coeff, sup = sample_and_dct(f, degs, a, b, retSupNorm=True)
            chunk = np.average(np.abs(coeff), axis=others)
            tol = absTol + sup * relTol
            guess *= 2
            if not started_converging(chunk, tol):
                continue
            degs[axis] = guess + 1
            coeff2, sup2 = sample_and_dct(f, degs, a, b, retSupNorm=True)
            tol = absTol + max(sup, sup2) * relTol
            if not has_converged(coeff, coeff2, tol):
                continue
            chunk = np.average(np.abs(coeff2), axis=others)
            deg, eps, rho = final_degree(chunk, tol)
            found[axis] = deg
            eps_list.append(eps)
            rho_list.append(rho)
            break
    return np.array(found), np.array(eps_list), np.array(rho_list)


def approx_error(degs, epsilons, rhos):
    """getApproxError :313-356 -- geometric tails over the 2^dim - 1 truncated index classes."""
    total = 0
    for mask in itertools.product(range(2), repeat=len(degs)):
        if np.sum(mask) == 0:
            continue
        weight = 1
        eps = 0
        for i, beyond in enumerate(mask):
            if beyond:
                weight /= (rhos[i] - 1)
                eps = max(eps, epsilons[i])
            else:
                weight *= (degs[i] + 1)
        total += weight * eps
    return total


def cheb_approximate(f, a, b, relApproxTol=1e-10):
    """chebApproximate :358-450 -> (coefficient tensor, error bound)."""
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
        f(mid) if isinstance(f, (MultiPower, MultiCheb)) else f(*mid)
    except TypeError:
        raise ValueError("Invalid input: length of the upper/lower bound lists must match the dimension of the function")
    degs, epsilons, rhos = chebyshev_degrees(f, a, b, relApproxTol)
    return sample_and_dct(f, degs, a, b), approx_error(degs, epsilons, rhos)


chebApproximate = cheb_approximate   # the reference's name
