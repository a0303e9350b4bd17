import sys, warnings
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
warnings.simplefilter("ignore")
from rootfinding_b200 import ChebyshevSubdivisionSolver as S
import helpers as H
for (deg,N) in [(12,6),(30,3)]:
    systems = H.seeded_systems(2, deg, N)
    errs = [np.full(2, 2.**-52)]*N
    sess, boxes, worst = H.check_batch_against_oracle(S.solve_batch, systems, errs)
    print(deg, N, boxes, sess.stats.boxes, worst)
