"""Where the end-to-end time of solve_batch goes (host buffers in, roots out): wall clock per phase + cProfile top.
usage: python scripts/e2e_profile.py [batch] [degree]"""
import cProfile, pstats, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rootfinding_b200 import ChebyshevSubdivisionSolver as S
from rootfinding_b200.runtime import get_engine
from rootfinding_b200.workloads import random_systems

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
deg = int(sys.argv[2]) if len(sys.argv) > 2 else 50
systems, errs = random_systems(2, deg, batch, seed=2)
eng = get_engine()
for _ in range(3):
    S.solve_batch(systems, errs)
torch.cuda.synchronize()
for rep in range(3):
    h0, d0 = eng.mem.h2d_bytes, eng.mem.d2h_bytes
    t0 = time.perf_counter()
    roots, sess = S.solve_batch(systems, errs, return_session=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"rep {rep}: {dt*1e3:.1f} ms, boxes {sess.stats.boxes}, {sess.stats.boxes/dt/1e6:.1f} M boxes/s, h2d {eng.mem.h2d_bytes-h0} d2h {eng.mem.d2h_bytes-d0}"
          f" sparse {sorted(getattr(sess, 'sparse_systems', []))}")
    del sess
pr = cProfile.Profile()
pr.enable()
roots = S.solve_batch(systems, errs)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
