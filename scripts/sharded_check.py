"""Under torchrun (NCCL): solve_sharded against the single-GPU solve_batch on the same inputs -- array_equal roots, equal
box counts -- for the test cases of tests/helpers.sharded_cases plus one random degree-50 system and notebook Demo 4.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 scripts/sharded_check.py
"""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
warnings.simplefilter("ignore")
import helpers as H, cases_r2 as cz
from rootfinding_b200 import ChebyshevSubdivisionSolver as S, ChebyshevApproximator as CA
from rootfinding_b200.distributed import solve_sharded

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cases = H.sharded_cases() + [("one50", H.seeded_systems(2, 50, 1)), ("eight100", H.seeded_systems(2, 100, 8))]
funcs, a, b = cz.NOTEBOOK["demo4"]
demo = [CA.chebApproximate(f, np.array(a, float), np.array(b, float)) for f in funcs]
cases.append(("demo4", [[demo[0][0], demo[1][0]]]))
demo_errs = np.array([demo[0][1], demo[1][1]])
ok = True
for name, systems in cases:
    errs = [demo_errs] if name == "demo4" else [np.full(2, 2.0 ** -52)] * len(systems)
    for every in (1, 2):
        torch.cuda.synchronize(); dist.barrier(); t0 = time.time()
        roots, info = solve_sharded(systems, errs, every=every, min_move=64, tolerance=1.1, return_stats=True)
        torch.cuda.synchronize(); dt = time.time() - t0
        if rank == 0:
            t1 = time.time()
            want, sess = S.solve_batch(systems, errs, return_session=True)
            torch.cuda.synchronize(); ds = time.time() - t1
            # "corners" has roots ON subdivision-grid corners (ill conditioned by construction): the merged re-solve sees sums
            # reduced by 16- or 32-lane teams depending on how many boxes a round holds -> last-bit differences
            same = all((np.array_equal(np.asarray(x).reshape(-1, 2), np.asarray(y).reshape(-1, 2)) if name != "corners" else
                        (np.shape(x) == np.shape(y) and (len(x) == 0 or H.match_points(x, y, 1e-9) < 1e-9))) for x, y in zip(roots, want))
            ok = ok and same and info["boxes"] == sess.stats.boxes
            print(f"{name:9s} every={every} world={world}: roots {sum(len(r) for r in roots)} equal={same} boxes {info['boxes']} (single {sess.stats.boxes}) "
                  f"per rank {info['boxes_per_rank']} moved {info['boxes_moved']} boxes / {info['bytes_moved']} B in {info['rebalances']} rebalances; "
                  f"{dt * 1e3:.1f} ms sharded vs {ds * 1e3:.1f} ms single", flush=True)
        dist.barrier()
if rank == 0:
    print("ALL EQUAL" if ok else "MISMATCH")
dist.destroy_process_group()
