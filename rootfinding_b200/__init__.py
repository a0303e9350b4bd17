"""rootfinding_b200 -- a B200-native drop-in for the hot path of YRoots (tylerjarvis/RootFinding).

    import rootfinding_b200 as yr
    roots = yr.solve([f, g], a, b)                      # yroots.solve
    yr.MultiCheb(coeff), yr.MultiPower(coeff)           # yroots.MultiCheb / MultiPower

The public surface mirrors /root/reference/yroots/__init__.py:4-6; the arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI of include/yroots_b200.h.  There is no CPU
fallback: using the solver without the CUDA library and a device raises.
"""
from .polynomial import MultiCheb, MultiPower          # noqa: F401
from .Combined_Solver import solve                     # noqa: F401
from . import ChebyshevApproximator, ChebyshevSubdivisionSolver, QuadraticCheck, Combined_Solver  # noqa: F401
from .ChebyshevSubdivisionSolver import solve_batch    # noqa: F401
from .ChebyshevApproximator import chebApproximate as approximate   # noqa: F401  (docs of the reference name it so)
from .ChebyshevApproximator import device_function     # noqa: F401  torch-vectorised callables are sampled on the device

__all__ = ["solve", "solve_batch", "approximate", "device_function", "MultiCheb", "MultiPower"]
