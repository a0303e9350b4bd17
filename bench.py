#!/usr/bin/env python
"""bench.py -- subdivision boxes/sec on batched random dense 2-D degree-50 Chebyshev systems.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl own|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path (device frontier solve) over one batch of B synthetic systems per
GPU: `rand_coeffs(dim=2, deg=50, B, 'randn')` of the reference's tests/random_tests.py:5-19, solved as
solveChebyshevSubdivision(Ms, errors=2^-52).  Weak scaling: every rank owns its own B systems.

  value      boxes/s (boxes = invocations of the reference's solvePolyRecursive, counted on the
             device), inputs resident in HBM, CUDA events around the level loop, max over ranks
  e2e        the same metric through the public API rootfinding_b200.solve_batch with HOST buffers:
             pack + pinned H2D + solve + D2H of the result records + host post-pass, per step
  roofline   for the dominant kernel family (the frontier pipeline's tensor ops f2_tensor_warp /
             f2_tensor_cta: restriction + subdivision of every box): algorithmic bytes (16 B per
             coefficient of every box at entry) / their summed device time, from CUDA
             events recorded inside yr_frontier_solve around the tensor launches of every round
  cpu_baseline  the oracle port of the reference solver on the host cores (one process per core)
`--impl reference` times that CPU arm alone on the same workload (rank 0 only).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

DIM, DEG = 2, 50
print_line = lambda obj: print(json.dumps(obj))
METRIC = "subdivision boxes/sec"
UNIT = "boxes/s"


def workload(batch, seed):
    from rootfinding_b200.workloads import rand_coeffs  # restates the reference generator tests/random_tests.py:5-19
    np.random.seed(seed)
    arr = rand_coeffs(DIM, DEG, batch, "randn")
    systems = [[np.ascontiguousarray(m) for m in arr[s]] for s in range(batch)]
    return systems, [np.full(DIM, 2.0 ** -52)] * batch


def measured_traffic():
    """DRAM bytes per tensor-kernel launch from the committed ncu launch list of the newest round (profiles/rNN_traffic.json:
    dram__bytes_read + dram__bytes_write summed over every tensor launch of two solves of this workload / number of launches)."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), reverse=True):
        try:
            with open(path) as fh:
                t = json.load(fh)
            return float(t["dram_bytes_per_launch"]), float(t["traffic_over_algorithmic"]), os.path.relpath(path, ROOT)
        except Exception:
            continue
    return None, None, None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_arm(steps, warmup, cores=None):
    from oracle.cpu_bench import CpuArm
    arm = CpuArm(DIM, DEG, cores=cores, per_worker=1)
    for _ in range(warmup):
        arm.step()
    boxes = roots = 0
    wall = 0.0
    for _ in range(steps):
        b, r, w = arm.step()
        boxes, roots, wall = boxes + b, roots + r, wall + w
    arm.close()
    return {"value": boxes / wall, "unit": UNIT, "cores": arm.cores, "kind": "port",
            "sample": f"{steps} x {arm.cores} random 2-D degree-{DEG} systems (one per core per step), "
                      f"{boxes} boxes in {wall:.1f} s; oracle/ port of the reference solver, pinned bit-for-bit to it"}, wall, steps


def run_reference(args, rank):
    if rank != 0:
        return
    base, wall, steps = cpu_arm(args.steps, min(args.warmup, 1))
    # the CPU arm solves a BOUNDED SAMPLE of the workload per step (one system of the same generator per host core, a few
    # seconds of CPU work); boxes/s is a per-box rate, so it compares with the GPU arm's rate on the full batch
    cfg = workload_config(args, args.batch)            # the workload definition of the GPU arm (what the ratio refers to)
    line = {"impl": "reference", "systems_per_step": base["cores"] * 1,
            "step_note": (f"a step of this arm solves {base['cores']} systems (one per host core) of the workload's generator, "
                          f"not the {args.batch} systems per GPU of config: ms_per_step belongs to that smaller step; "
                          "boxes/s is per box and compares directly"), "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg, "cpu_baseline": base, "gpu_launches": 0,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print_line(line)


def workload_config(args, systems_per_step):
    return {"workload": f"batched random dense 2-D Chebyshev systems, degree {DEG} (BASELINE.json configs[2], headline deg 50): "
                        f"rand_coeffs(2,{DEG},B,'randn'), errors 2^-52",
            "systems_per_gpu_per_step": systems_per_step, "dim": DIM, "degree": DEG,
            "l2": "frontier traffic per step >> 126 MB L2; 256 MB scratch written between steps",
            "arithmetic": "fma" if args.fma else "separately rounded mul/add (restrictions bit-identical to the reference)"}


def run_strong(args, rank, world, eng):
    """Box-level multi-GPU mode: one fixed batch, its frontier sharded over the ranks (rootfinding_b200.distributed.
    solve_sharded: queue lengths all-gathered every `--every` rounds, surplus boxes shipped with NCCL send/recv).  Timed with
    CUDA events on each rank's stream from the first level-0 launch to the arrival of the broadcast roots, max over ranks."""
    import torch
    import torch.distributed as dist
    from rootfinding_b200.distributed import solve_sharded
    from rootfinding_b200.workloads import rand_coeffs
    np.random.seed(2)
    arr = rand_coeffs(DIM, args.strong_degree, args.strong_systems, "randn")
    systems = [[np.ascontiguousarray(m) for m in arr[s]] for s in range(args.strong_systems)]
    errs = [np.full(DIM, 2.0 ** -52)] * len(systems)
    total_ms, boxes, info, solve_ms, asm_ms = 0.0, 0, None, 0.0, 0.0
    for it in range(args.warmup + args.steps):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        roots, info = solve_sharded(systems, errs, every=args.every, use_fma=bool(args.fma), return_stats=True, start=args.strong_start)
        e1.record()
        torch.cuda.synchronize(); dist.barrier()
        if it >= args.warmup:
            total_ms += e0.elapsed_time(e1)
            boxes += info["boxes"]
            solve_ms += info["solve_ms"]; asm_ms += info["assemble_ms"]
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t[0])
        # value: the sharded search itself (H2D of the coefficients, level-0 boxes, every frontier round and rebalance; slowest
        # rank, host clock bracketed by stream synchronisations); e2e: the whole call, CUDA events, incl. the record gather
        # to rank 0, the serial tree bookkeeping there and the broadcast of the roots
        print_line({"metric": METRIC, "value": boxes / (solve_ms * 1e-3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": solve_ms / args.steps,
                    "e2e": {"value": boxes / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / args.steps, "assemble_ms_per_step": asm_ms / args.steps}, "higher_is_better": True, "scaling": "strong",
                    "vs_baseline": None, "dtype": "f64", "data": "synthetic", "mode": "strong",
                    "config": {"workload": f"ONE batch of {args.strong_systems} random dense 2-D degree-{args.strong_degree} systems "
                                           f"(rand_coeffs, seed 2), frontier sharded over the GPUs, rebalanced every {args.every} rounds, start={args.strong_start}",
                               "dim": DIM, "degree": args.strong_degree, "systems_total": args.strong_systems},
                    "detail": {"roots": int(sum(len(r) for r in roots)), "boxes_per_step": info["boxes"], "boxes_per_rank": info["boxes_per_rank"],
                               "boxes_moved_per_step": info["boxes_moved"], "bytes_moved_per_step": info["bytes_moved"],
                               "rebalances_per_step": info["rebalances"]}})
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("YR_BENCH_BATCH", "1024")))
    ap.add_argument("--fma", type=int, default=int(os.environ.get("YR_BENCH_FMA", "0")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="weak", choices=["weak", "strong"],
                    help="strong: a FIXED batch of few large systems, its frontier sharded over the GPUs with NCCL "
                         "load rebalancing (documented side measurement, not the headline)")
    ap.add_argument("--strong-systems", type=int, default=8)
    ap.add_argument("--strong-degree", type=int, default=100)
    ap.add_argument("--every", type=int, default=2, help="strong mode: rounds between two rebalancing points")
    ap.add_argument("--strong-start", default="round_robin", choices=["round_robin", "rank0"],
                    help="strong mode: who starts the systems (rank0: the rebalancer has to spread the whole frontier)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warnings.simplefilter("ignore")
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    global print_line
    print_line = lambda obj: (real_stdout.write(json.dumps(obj) + "\n"), real_stdout.flush())
    if args.impl == "reference":
        return run_reference(args, rank)

    base = None

    import torch
    import torch.distributed as dist
    from rootfinding_b200 import ChebyshevSubdivisionSolver as S, postpass
    from rootfinding_b200.runtime import get_engine
    from rootfinding_b200.distributed import all_gather_roots

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    elif args.mode == "strong":
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{29400 + os.getpid() % 500}", rank=0, world_size=1,
                                device_id=torch.device("cuda", local_rank))
    eng = get_engine()
    if args.mode == "strong":
        return run_strong(args, rank, world, eng)
    systems, errs = workload(args.batch, seed=2 + rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(timed):
        """Inputs resident in HBM; timed region = level-0 box creation + every level launch."""
        sess = eng.session(systems, errs, use_fma=args.fma)
        sess.time_kernels = timed
        flush.fill_(1)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sess.run_jobs(sess.unit_jobs())
        e1.record()
        barrier()
        return sess, e0.elapsed_time(e1)

    gather_ms = []                                             # host time of the final root all-gather per e2e step (this rank)

    def api_step():
        """Public API with host buffers: pack, pinned H2D, solve, D2H, post-pass, root all-gather."""
        flush.fill_(1)
        barrier()
        h0, d0 = eng.mem.h2d_bytes, eng.mem.d2h_bytes
        t0 = time.perf_counter()
        roots, sess = S.solve_batch(systems, errs, use_fma=bool(args.fma), return_session=True)
        t1 = time.perf_counter()
        if world > 1:
            all_gather_roots(roots, np.arange(rank, world * len(systems), world), world * len(systems), DIM)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        gather_ms.append((time.perf_counter() - t1) * 1e3)
        barrier()
        # only numbers leave this function: a session that stays referenced keeps its page-locked record buffers
        return sess.stats.boxes, dt, eng.mem.h2d_bytes - h0, eng.mem.d2h_bytes - d0, sum(len(r) for r in roots)

    for _ in range(args.warmup):
        device_step(False)
    sampler = ClockSampler(local_rank)
    if rank == 0:                                              # one sampler per job: rank 0 prints the line
        sampler.start()
    total_ms, boxes, launches, kernel_ms, kernel_bytes, scalar_ms, tensor_launches = 0.0, 0, 0, 0.0, 0, 0.0, 0
    for _ in range(args.steps):
        sess, ms = device_step(True)
        total_ms += ms
        boxes += sess.stats.boxes
        launches += sess.stats.launches
        kernel_ms += sum(k[0] for k in sess.kernel_ms)
        kernel_bytes += sum(k[1] for k in sess.kernel_ms)
        scalar_ms += getattr(sess, "scalar_ms", 0.0)
        stats = sess.stats
        tensor_launches += sess.stats.launches - getattr(sess.stats, "rounds", 0) - 2      # minus scalar steps, import, tables
    e2e_s, e2e_boxes, h2d, d2h, nroots = 0.0, 0, 0, 0, 0
    api_step()                                                 # warm the API path once
    for _ in range(args.steps):
        nb, dt, hb, db, nroots = api_step()
        e2e_s += dt
        e2e_boxes += nb
        h2d, d2h = hb, db
    sampler.stop_flag.set()
    if rank == 0:
        sampler.join(timeout=2)

    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        # bounded sample of the same workload on the host cores; after the GPU arms so that 16 busy worker
        # processes never share the machine with the (latency-sensitive, single-threaded) round loop
        base, _, _ = cpu_arm(steps=2, warmup=0)

    t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([boxes, e2e_boxes, launches], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)               # slowest rank defines the step
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    total_ms_max, e2e_s_max = float(t[0]), float(t[1])
    boxes_all, e2e_boxes_all, launches_all = float(cnt[0]), float(cnt[1]), int(cnt[2])
    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = kernel_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
        traffic, traffic_ratio, traffic_file = measured_traffic()
        line = {"metric": METRIC, "value": boxes_all / (total_ms_max * 1e-3), "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args, args.batch),
                "e2e": {"value": e2e_boxes_all / e2e_s_max, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "solves_per_s": args.batch * world * args.steps / e2e_s_max, "roots_per_step_rank0": int(nroots),
                        "root_allgather_ms_per_step_rank0": (sum(gather_ms[1:]) / max(1, len(gather_ms) - 1)) if world > 1 else 0.0},
                "gpu_launches": launches_all,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "traffic_over_algorithmic": traffic_ratio,
                             "traffic_source": f"{traffic_file} (ncu dram__bytes_read.sum + dram__bytes_write.sum per tensor launch, same workload)",
                             "algorithmic_bytes_per_launch": kernel_bytes / max(tensor_launches, 1),
                             "kernel": "f2_tensor_warp / f2_tensor_cta (restrict + subdivide launches of every round of the timed steps, rank 0)",
                             "fp64_frac_of_measured_dmul_dadd_peak": (2.0 * stats.restrict_macs / 1e12 / max(kernel_ms / args.steps * 1e-3, 1e-12)) / 18.56,
                             "scalar_kernel_ms_per_step": scalar_ms / args.steps,
                             "peak_source": peak_src,
                             "bytes_per_box": 16.0 * stats.entry_coeffs / max(stats.boxes, 1),
                             "kernel_share_of_step": kernel_ms / max(total_ms, 1e-9)},
                "clocks": sampler.summary(),
                "detail": {"boxes_per_step_rank0": stats.boxes, "rounds": getattr(stats, "rounds", stats.levels), "retries": stats.retries,
                           "zooms": stats.zooms, "subdivisions": stats.subdivisions,
                           "restrict_gflops_per_s": 2.0 * stats.restrict_macs / 1e9 / (total_ms / args.steps * 1e-3),
                           "build": eng.lib.yr_build_info().decode(),
                           "boxes_vs_reference": "own box count == the reference's on the same inputs (1 792 seeded systems, "
                                                 "profiles/r01_parity_stress.txt; tests/test_gpu_parity.py), so boxes/s is also "
                                                 "reference-equivalent boxes/s (SURVEY 8d)"}}
        if base is not None:
            line["cpu_baseline"] = base
        print_line(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
