"""ORACLE -- test infrastructure only.  CPU arm of bench.py: the oracle port of the reference's
subdivision solver timed on the host cores (one process per core, systems dealt round-robin)."""
import os
import time
import warnings

import numpy as np


def _workload(dim, deg, count, seed):
    from oracle.gen_golden import rand_coeffs
    np.random.seed(seed)
    batch = rand_coeffs(dim, deg, count, "randn")
    return [[np.ascontiguousarray(m) for m in batch[s]] for s in range(count)]


def _solve_some(args):
    dim, deg, seed, count = args
    from oracle import subdivision as osub
    warnings.simplefilter("ignore")
    systems = _workload(dim, deg, count, seed)
    boxes = roots = 0
    t0 = time.perf_counter()
    for Ms in systems:
        trace = osub.Trace()
        r = osub.solve_chebyshev_subdivision(Ms, np.full(dim, 2.0 ** -52), trace=trace)
        boxes += trace.boxes
        roots += len(r)
    return boxes, roots, time.perf_counter() - t0


class CpuArm:
    """A pool of worker processes; `step()` solves `per_worker` fresh systems on every core."""

    def __init__(self, dim, deg, cores=None, per_worker=1):
        import multiprocessing as mp
        self.dim, self.deg = dim, deg
        self.cores = cores or os.cpu_count() or 1
        self.per_worker = per_worker
        self.pool = mp.get_context("spawn").Pool(self.cores)
        self.seed = 1000
        self.pool.map(_solve_some, [(dim, min(deg, 6), 0, 1)] * self.cores)        # import + page in, untimed

    def step(self):
        jobs = [(self.dim, self.deg, self.seed + k, self.per_worker) for k in range(self.cores)]
        self.seed += self.cores
        t0 = time.perf_counter()
        out = self.pool.map(_solve_some, jobs)
        wall = time.perf_counter() - t0
        return sum(o[0] for o in out), sum(o[1] for o in out), wall

    def close(self):
        self.pool.close()
        self.pool.join()
