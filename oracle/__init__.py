"""ORACLE -- TEST INFRASTRUCTURE ONLY.

A CPU restatement (numpy + one small C file) of the reference's hot path
(tylerjarvis/RootFinding: yroots/Combined_Solver.py, ChebyshevApproximator.py,
ChebyshevSubdivisionSolver.py, QuadraticCheck.py, polynomial.py).  It exists to
check the CUDA implementation and to serve as the timed CPU baseline.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import it.  The product package
(rootfinding_b200/) never does, and fails loudly without its CUDA library.

Parity status: PINNED -- oracle/gen_golden.py runs the unmodified reference
(imported from /root/reference with a one-line `np.product = np.prod` shim for
numpy >= 2) in the build container and stores its inputs/outputs under
tests/golden/; tests/test_oracle_vs_golden.py checks this oracle against them,
together with the reference's own golden roots (Polished_results/,
Chebfun_results/, copied as data fixtures).
"""
from . import kernels, quadratic, subdivision, polynomial, approximator, combined  # noqa: F401
from .combined import solve  # noqa: F401
from .polynomial import MultiCheb, MultiPower  # noqa: F401
