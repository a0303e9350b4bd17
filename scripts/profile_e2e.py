"""Host-side profile of the public API path (solve_batch with host buffers) on a batch of random 2-D systems."""
import cProfile, os, pstats, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
warnings.simplefilter("ignore")
from rootfinding_b200 import ChebyshevSubdivisionSolver as S
from rootfinding_b200.workloads import rand_coeffs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
deg = int(sys.argv[2]) if len(sys.argv) > 2 else 50
np.random.seed(2)
arr = rand_coeffs(2, deg, B)
systems = [[np.ascontiguousarray(m) for m in arr[s]] for s in range(B)]
errs = [np.full(2, 2.0 ** -52)] * B
for rep in range(2):
    t = time.perf_counter(); roots = S.solve_batch(systems, errs); torch.cuda.synchronize(); print("solve_batch %.1f ms" % ((time.perf_counter() - t) * 1e3))
pr = cProfile.Profile(); pr.enable()
roots = S.solve_batch(systems, errs); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
