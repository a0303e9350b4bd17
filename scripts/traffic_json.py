"""profiles/rNN_traffic.json from an ncu launch list (time + DRAM bytes per launch) and the bench line of the same workload.

    python scripts/traffic_json.py LAUNCHES.csv BENCH.json STEPS_IN_LIST KERNEL_REGEX > profiles/rNN_traffic.json
"""
import csv, json, re, sys
path, bench_path, steps, pattern = sys.argv[1], sys.argv[2], int(sys.argv[3]), re.compile(sys.argv[4])
with open(path) as fh:
    lines = [l for l in fh if l.startswith('"')]
byid = {}
for r in csv.DictReader(lines):
    if not pattern.search(r["Kernel Name"]):
        continue
    v, unit, m = float(r["Metric Value"].replace(",", "")), r["Metric Unit"].lower(), r["Metric Name"]
    if m.startswith("dram__bytes"):
        v *= {"k": 1e3, "m": 1e6, "g": 1e9}.get(unit[:1], 1.0) if unit != "byte" else 1.0
    elif m.startswith("gpu__time"):
        v *= {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}[unit]
    byid.setdefault(int(r["ID"]), {})[m] = v
n = len(byid)
dram = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in byid.values())
ms = sum(d.get("gpu__time_duration.sum", 0) for d in byid.values())
bench = json.loads(open(bench_path).read().strip().splitlines()[-1])
alg_step = bench["roofline"]["bytes_per_box"] * bench["detail"]["boxes_per_step_rank0"]
out = {"source": f"{path} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum, {steps} solves of the bench workload)",
       "kernel_family": pattern.pattern, "launches_per_step": n / steps, "dram_bytes_per_step": dram / steps,
       "dram_bytes_per_launch": dram / max(n, 1), "algorithmic_bytes_per_step": alg_step,
       "algorithmic_bytes_per_launch": alg_step / (n / steps), "traffic_over_algorithmic": dram / steps / alg_step,
       "ncu_serialised_ms_per_step": ms / steps}
print(json.dumps(out, indent=1))
