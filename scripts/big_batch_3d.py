"""Large 3-D / 4-D batches through solve_batch (level pipeline; exercises the capacity growth and the halving fallback)."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
warnings.simplefilter("ignore")
from rootfinding_b200 import ChebyshevSubdivisionSolver as S
from rootfinding_b200.workloads import random_systems
for dim, deg, B in [(3, 10, 512), (4, 4, 512)]:
    systems, errs = random_systems(dim, deg, B)
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        roots, sess = S.solve_batch(systems, errs, return_session=True)
        torch.cuda.synchronize(); dt = time.time() - t
        chunks = getattr(sess, "chunk_sessions", [sess])
        boxes = sum(c.stats.boxes for c in chunks)
        print(f"dim {dim} deg {deg} B {B} rep {rep}: {boxes} boxes, {sum(len(r) for r in roots)} roots, {dt*1e3:.0f} ms -> {boxes/dt:.3e} boxes/s, "
              f"{len(chunks)} chunk(s), retries {sum(c.stats.retries for c in chunks)}, peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
        del roots, sess, chunks
