"""Latency of the drop-in yr.solve (opaque numpy callables, host-sampled like the reference) per Chebfun2 system,
next to the CPU oracle port on the same machine.  Second call of each (no JIT / first-touch)."""
import os, sys, time, warnings, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
warnings.simplefilter("ignore")
import chebfun_cases as cc
import rootfinding_b200 as yr
from oracle import combined as ocomb
tids = sys.argv[1].split(",") if len(sys.argv) > 1 else ["1.1", "2.5", "7.4", "8.2", "10.1"]
for tid in tids:
    funcs, a, b = cc.system(tid)
    out = []
    for name, fn in (("b200", yr.solve), ("oracle", ocomb.solve)):
        fn(funcs, a, b)
        t = time.perf_counter(); r = fn(funcs, a, b); dt = time.perf_counter() - t
        out.append(f"{name} {dt*1e3:8.1f} ms ({len(r)} roots)")
    print(tid, " | ".join(out), flush=True)
if os.environ.get("YR_PROFILE"):
    funcs, a, b = cc.system(os.environ["YR_PROFILE"])
    pr = cProfile.Profile(); pr.enable(); yr.solve(funcs, a, b); pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)
