"""CPU tier: the N>1 path (system sharding + final root all-gather) with world_size 2 over gloo.
Each rank drives the kernel-source emulator; the result must equal the single-process solve."""
import os
import subprocess
import sys
import tempfile
import textwrap

import numpy as np

import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, warnings
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import numpy as np
    import torch.distributed as dist
    warnings.simplefilter("ignore")
    from emul.backend import emulated_engine
    from rootfinding_b200.distributed import solve_batch_sharded, shard
    import helpers as H
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    systems = H.seeded_systems(2, 8, 5)
    errs = [np.full(2, 2. ** -52)] * 5
    roots = solve_batch_sharded(systems, errs, engine=emulated_engine())
    np.savez({out!r} + sys.argv[1] + ".npz", **{{f"r{{i}}": r for i, r in enumerate(roots)}})
    dist.destroy_process_group()
""")


def test_world_size_two_gloo():
    from emul.backend import emulated_engine
    from rootfinding_b200 import ChebyshevSubdivisionSolver as S
    emulated_engine()                                                  # build the emulator once, before the ranks start
    with tempfile.TemporaryDirectory() as tmp:
        script = os.path.join(tmp, "worker.py")
        out = os.path.join(tmp, "roots_rank")
        with open(script, "w") as fh:
            fh.write(WORKER.format(root=ROOT, port=29533, out=out))
        procs = [subprocess.Popen([sys.executable, script, str(r)]) for r in range(2)]
        for p in procs:
            assert p.wait(timeout=600) == 0
        systems = H.seeded_systems(2, 8, 5)
        errs = [np.full(2, 2. ** -52)] * 5
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = S.solve_batch(systems, errs, engine=emulated_engine())
        for rank in range(2):
            z = np.load(out + f"{rank}.npz")
            for i, w in enumerate(want):
                assert np.array_equal(z[f"r{i}"], w), (rank, i)       # independent of the number of ranks


SHARD_WORKER = textwrap.dedent("""
    import os, sys, warnings
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import numpy as np
    import torch.distributed as dist
    warnings.simplefilter("ignore")
    from emul.backend import emulated_engine
    from rootfinding_b200.distributed import solve_sharded
    import helpers as H
    rank, world = int(sys.argv[1]), int(sys.argv[2])
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    for name, systems in H.sharded_cases():
        errs = [np.full(2, 2. ** -52)] * len(systems)
        roots, info = solve_sharded(systems, errs, engine=emulated_engine(), every=1, min_move=4, tolerance=1.05, return_stats=True)
        np.savez({out!r} + name + "_" + sys.argv[1] + ".npz", boxes=info["boxes"], moved=info["boxes_moved"], rebalances=info["rebalances"],
                 per_rank=np.array(info["boxes_per_rank"]), **{{f"r{{i}}": r for i, r in enumerate(roots)}})
    dist.destroy_process_group()
""")


def test_one_frontier_sharded_over_two_ranks_gloo():
    """Box-level mode: ONE system's frontier spread over two ranks with periodic rebalancing
    (yr_frontier_export / _import through gloo).  Roots are array_equal to the single-process solve, the total number of
    solvePolyRecursive invocations is the same, both ranks did real work, and boxes did move.  Cases: one random degree-30
    system; a batch with roots on subdivision-grid corners (exterior boxes of DIFFERENT ranks are merged on rank 0); one
    system with an extent above the frontier pipeline's limit (its first levels run on the level pipeline of its rank)."""
    from emul.backend import emulated_engine
    from rootfinding_b200 import ChebyshevSubdivisionSolver as S
    emulated_engine()
    with tempfile.TemporaryDirectory() as tmp:
        script = os.path.join(tmp, "worker.py")
        out = os.path.join(tmp, "shard_")
        with open(script, "w") as fh:
            fh.write(SHARD_WORKER.format(root=ROOT, port=29537, out=out))
        procs = [subprocess.Popen([sys.executable, script, str(r), "2"]) for r in range(2)]
        for p in procs:
            assert p.wait(timeout=900) == 0
        import warnings
        for name, systems in H.sharded_cases():
            errs = [np.full(2, 2. ** -52)] * len(systems)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                want, sess = S.solve_batch(systems, errs, engine=emulated_engine(), return_session=True)
            for rank in range(2):
                z = np.load(out + f"{name}_{rank}.npz")
                for i, w in enumerate(want):
                    assert np.array_equal(z[f"r{i}"], w), (name, rank, i)
                assert int(z["boxes"]) == sess.stats.boxes, (name, int(z["boxes"]), sess.stats.boxes)
                assert int(z["moved"]) > 0 and int(z["rebalances"]) > 0
                assert np.all(z["per_rank"] > 0.2 * sess.stats.boxes)


def test_rebalance_plan_is_deterministic_and_conserving():
    from rootfinding_b200.distributed import rebalance_plan
    rs = np.random.RandomState(0)
    for world in (2, 3, 4, 8):
        for _ in range(50):
            counts = rs.randint(0, 5000, size=world)
            plan = rebalance_plan(counts, min_move=16, tolerance=1.1)
            after = counts.astype(np.int64).copy()
            for src, dst, n in plan:
                assert n >= 16 and src != dst
                after[src] -= n; after[dst] += n
            assert after.sum() == counts.sum() and np.all(after >= 0)
            if plan:
                assert after.max() <= counts.max()
                assert after.max() - after.min() <= max(16 * world, counts.max() - counts.min())
            assert plan == rebalance_plan(list(counts), min_move=16, tolerance=1.1)
    assert rebalance_plan([10, 0], min_move=256) == []            # too small to be worth moving
    assert rebalance_plan([1000, 990], min_move=16, tolerance=1.25) == []


def test_shard_is_a_partition():
    from rootfinding_b200.distributed import shard
    for n in (0, 1, 7, 64):
        for world in (1, 2, 4, 8):
            parts = [shard(n, r, world) for r in range(world)]
            assert sorted(np.concatenate(parts).tolist()) == list(range(n))
