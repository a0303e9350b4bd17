"""Per-kernel shares of an ncu launch list (gpu__time_duration.sum [+ dram__bytes_read/write.sum]) CSV.

    python scripts/launch_shares.py LAUNCHES.csv [first_id [last_id]]
"""
import csv, sys, collections, re
path = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
rows = []
with open(path) as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.DictReader(lines)
per = collections.defaultdict(lambda: collections.defaultdict(float))
ids = collections.defaultdict(set)
for r in rd:
    i = int(r["ID"])
    if i < lo or i > hi:
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = name.replace("<unnamed>::", "").replace("yr_f2_backend::", "")
    if len(name) > 60: name = name[:60]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    m = r["Metric Name"]
    if m.startswith("dram__bytes") and unit.lower().startswith("k"): v *= 1e3
    if m.startswith("dram__bytes") and unit.lower().startswith("m"): v *= 1e6
    if m.startswith("dram__bytes") and unit.lower().startswith("g"): v *= 1e9
    if m.startswith("gpu__time") and unit == "us": v *= 1e3
    if m.startswith("gpu__time") and unit == "ms": v *= 1e6
    per[name][m] += v
    ids[name].add(i)
tot = sum(d["gpu__time_duration.sum"] for d in per.values())
print(f"{'kernel':60s} {'launches':>8s} {'time ms':>10s} {'share':>7s} {'dram rd GB':>11s} {'dram wr GB':>11s} {'GB/s':>8s}")
for name, d in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    t = d["gpu__time_duration.sum"]
    rdb, wrb = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
    print(f"{name:60s} {len(ids[name]):8d} {t/1e6:10.3f} {100*t/tot:6.1f}% {rdb/1e9:11.3f} {wrb/1e9:11.3f} {(rdb+wrb)/max(t,1):8.1f}")
print(f"total {tot/1e6:.3f} ms")
