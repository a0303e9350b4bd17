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
import cases_r2 as cz
from rootfinding_b200 import ChebyshevSubdivisionSolver as S, ChebyshevApproximator as CA
tids = sys.argv[1].split(",") if len(sys.argv) > 1 else ["1.1", "2.5", "7.4", "8.2", "10.1", "demo1", "demo2", "demo4"]
calls = {"solve_batch": 0, "approx": 0, "regions": 0}
_sb, _ca = S.solve_batch, CA.chebApproximate
def counted_sb(systems, *a, **k):
    calls["solve_batch"] += 1; calls["regions"] += len(systems)
    return _sb(systems, *a, **k)
def counted_ca(*a, **k):
    calls["approx"] += 1
    return _ca(*a, **k)
S.solve_batch, CA.chebApproximate = counted_sb, counted_ca
for tid in tids:
    funcs, a, b = cz.NOTEBOOK[tid] if tid in cz.NOTEBOOK else cc.system(tid)
    out = []
    for name, fn in (("b200", yr.solve), ("oracle", ocomb.solve)):
        fn(funcs, a, b)
        for k in calls: calls[k] = 0
        t = time.perf_counter(); r = fn(funcs, a, b); dt = time.perf_counter() - t
        extra = f", {calls['solve_batch']} device solves for {calls['regions']} regions, {calls['approx']} approximations" if name == "b200" else ""
        out.append(f"{name} {dt*1e3:8.1f} ms ({len(r)} roots{extra})")
    print(tid, " | ".join(out), flush=True)
if os.environ.get("YR_PROFILE"):
    funcs, a, b = cc.system(os.environ["YR_PROFILE"])
    pr = cProfile.Profile(); pr.enable(); yr.solve(funcs, a, b); pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)
