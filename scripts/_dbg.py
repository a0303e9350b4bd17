import sys, warnings, numpy as np, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
warnings.simplefilter("ignore")
import helpers as H, cases_r2 as cz
import rootfinding_b200 as yr
S = yr.ChebyshevSubdivisionSolver
for dim, deg, N, kind in cz.RANDOM_PLAN:
    if dim not in (3, 4): continue
    systems = H.seeded_systems(dim, deg, N, kind)
    for s in range(N):
        try:
            got, sess = S.solve_batch([systems[s]], [np.full(dim, 2.**-52)], return_session=True)
            print(dim, deg, s, kind, "ok boxes", sess.stats.boxes, flush=True)
        except Exception as e:
            print(dim, deg, s, kind, "FAIL", str(e)[:200], flush=True)
