"""First GPU contact: parity of the CUDA solver against the oracle + a rough timing."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
warnings.simplefilter("ignore")
from rootfinding_b200 import ChebyshevSubdivisionSolver as S
from rootfinding_b200.runtime import get_engine
from oracle import subdivision as osub
from rootfinding_b200.workloads import rand_coeffs

print(torch.cuda.get_device_name(0))
eng = get_engine()
print(eng.lib.yr_build_info().decode(), "SMs", eng.sm_count, "smem/cta", eng.max_smem_cta, "smem/sm", eng.smem_per_sm)

def run(dim, deg, N, kind="randn", exact=False, check=True, use_fma=False):
    np.random.seed(2)
    batch = rand_coeffs(dim, deg, N, kind)
    systems = [[np.ascontiguousarray(m) for m in batch[s]] for s in range(N)]
    errs = [np.full(dim, 2. ** -52)] * N
    torch.cuda.synchronize(); t = time.time()
    got, sess = S.solve_batch(systems, errs, exact=exact, use_fma=use_fma, return_session=True)
    torch.cuda.synchronize(); tg = time.time() - t
    tot = 0; bad = 0; md_all = 0.0
    if check:
        for s in range(N):
            tr = osub.Trace()
            want = np.asarray(osub.solve_chebyshev_subdivision([m.copy() for m in systems[s]], errs[s].copy(), exact=exact, trace=tr)).reshape(-1, dim)
            tot += tr.boxes
            g = got[s]
            if want.shape != g.shape:
                bad += 1; print("   COUNT MISMATCH sys", s, want.shape, g.shape)
            elif want.size:
                md_all = max(md_all, np.abs(want[np.lexsort(want.T)] - g[np.lexsort(g.T)]).max())
    st = sess.stats
    print(f"dim{dim} deg{deg} N={N} {kind} exact={exact} fma={use_fma}: boxes oracle {tot} device {st.boxes} roots {sum(len(g) for g in got)} "
          f"maxdiff {md_all:.2e} mismatches {bad} levels {st.levels} launches {st.launches} retries {st.retries} time {tg:.3f}s "
          f"-> {st.boxes/tg:.3e} boxes/s")
    return sess

run(2, 5, 4)
run(1, 10, 3); run(2, 10, 3); run(2, 10, 3, exact=True); run(3, 3, 2); run(3, 6, 2); run(4, 2, 2); run(4, 3, 1)
run(2, 20, 3); run(2, 20, 3, use_fma=True)
run(2, 50, 1)
for N in (16, 128):
    s = run(2, 20, N, check=False)
s = run(2, 20, 128, check=False)
print(s.stats.per_level)
s = run(2, 50, 16, check=False)
print(s.stats.per_level)
