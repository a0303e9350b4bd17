"""Parity stress (checker script, not part of the product): many seeded random systems solved by the CUDA library in
one batch and by the CPU oracle on all host cores; compares per-system box counts, root counts and roots (1e-10).

    python scripts/parity_stress.py "dim,deg,N,kind,seed;..."
"""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np


def oracle_one(args):
    dim, Ms = args
    warnings.simplefilter("ignore")
    from oracle import subdivision as osub
    tr = osub.Trace()
    r = np.asarray(osub.solve_chebyshev_subdivision([m.copy() for m in Ms], np.full(dim, 2.0 ** -52), trace=tr)).reshape(-1, dim)
    return tr.boxes, r


def main():
    import multiprocessing as mp
    warnings.simplefilter("ignore")
    from rootfinding_b200.workloads import rand_coeffs
    from rootfinding_b200 import ChebyshevSubdivisionSolver as S
    spec = sys.argv[1] if len(sys.argv) > 1 else "2,20,128,randn,7;2,35,64,randn,8;2,50,32,randn,9;2,12,128,randint,10;3,5,32,randn,11;1,40,64,randn,12"
    pool = mp.get_context("spawn").Pool(os.cpu_count())
    bad = 0
    for item in spec.split(";"):
        dim, deg, N, kind, seed = item.split(",")
        dim, deg, N, seed = int(dim), int(deg), int(N), int(seed)
        np.random.seed(seed)
        arr = rand_coeffs(dim, deg, N, kind)
        systems = [[np.ascontiguousarray(m) for m in arr[s]] for s in range(N)]
        errs = [np.full(dim, 2.0 ** -52)] * N
        t = time.time()
        per = [S.solve_batch([Ms], [e], return_session=True) for Ms, e in zip(systems[:8], errs[:8])]     # per-system box counts
        roots, sess = S.solve_batch(systems, errs, return_session=True)
        tg = time.time() - t
        t = time.time()
        ora = pool.map(oracle_one, [(dim, Ms) for Ms in systems])
        tc = time.time() - t
        tot_o = sum(o[0] for o in ora)
        worst, nroots, mism = 0.0, 0, 0
        for s, (o, r) in enumerate(zip(ora, roots)):
            r = np.asarray(r).reshape(-1, dim)
            if len(r) != len(o[1]):
                mism += 1
                continue
            nroots += len(r)
            if len(r):
                a, b = o[1][np.lexsort(o[1].T[::-1])], r[np.lexsort(r.T[::-1])]
                d = np.max(np.abs(a - b))
                if d > 1e-10:       # sort order can differ for near-ties: fall back to nearest matching
                    d = max(np.min(np.max(np.abs(b - x), axis=1)) for x in a)
                worst = max(worst, d)
        box8 = sum(p[1].stats.boxes for p in per) == sum(o[0] for o in ora[:8])
        ok = (sess.stats.boxes == tot_o) and mism == 0 and worst < 1e-10 and box8
        bad += not ok
        print(f"dim {dim} deg {deg:3d} N {N:4d} {kind:7s}: boxes gpu {sess.stats.boxes} oracle {tot_o} | systems with other root count {mism} | "
              f"roots {nroots} worst |diff| {worst:.2e} | gpu {tg:.2f} s oracle {tc:.1f} s ({os.cpu_count()} cores) | {'OK' if ok else 'MISMATCH'}", flush=True)
    pool.close()
    print("parity stress:", "all OK" if not bad else f"{bad} configuration(s) differ")


if __name__ == "__main__":
    main()
