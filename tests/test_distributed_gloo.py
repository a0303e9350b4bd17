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


def test_shard_is_a_partition():
    from rootfinding_b200.distributed import shard
    for n in (0, 1, 7, 64):
        for world in (1, 2, 4, 8):
            parts = [shard(n, r, world) for r in range(world)]
            assert sorted(np.concatenate(parts).tolist()) == list(range(n))
