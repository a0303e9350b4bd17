"""Per-source-line profile from an ncu report: joins the SASS page of `ncu --page source --csv` with the
line table nvdisasm prints for the same kernel in the built object (-lineinfo).

    python scripts/ncu_by_line.py REPORT.ncu-rep KERNEL_SUBSTRING [launch_index] [object]
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, ksub = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
obj = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "rootfinding_b200", "csrc", "yr_f2_kernels.o")
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# split per kernel
blocks, cur = [], None
for row in csv.reader(io.StringIO(txt)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(row)
sel = [b for b in blocks if ksub in b["name"]]
blk = sel[which]
hdr = blk["rows"][0]
rows = blk["rows"][1:]
ia, ii, it, iss = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[0][ia], 16)

# line table from nvdisasm
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# mangled-name match: strip template punctuation from the ncu name
m0 = re.search(r"(\w+)<\(int\)(\d+), \(bool\)(\d)(?:, \(int\)(\d+))?>", blk["name"])
mkey = os.environ.get("YR_MANGLED")            # explicit substring of the mangled name (kernels with other template lists)
if mkey:
    key = mkey
elif m0:
    key = f"{m0.group(1)}ILi{m0.group(2)}ELb{m0.group(3)}E" + (f"Li{m0.group(4)}E" if m0.group(4) else "")
else:
    key = re.search(r"(\w+)[<(]", blk["name"]).group(1)
line_of, infn, curline = {}, False, ("?", 0)
for ln in dis.splitlines():
    if ln.startswith(".text.") and ln.rstrip().endswith(":"):
        infn = key in ln
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        curline = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (curline, m.group(2).strip())
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
tot_i = tot_s = 0
for r in rows:
    off = int(r[ia], 16) - base
    (f, l), _ = line_of.get(off, (("?", 0), ""))
    a = agg[(f, l)]
    n, t, s = int(r[ii] or 0), int(r[it] or 0), int(r[iss] or 0)
    a[0] += n; a[1] += t; a[2] += s
    for c in stall_cols:
        v = int(r[c] or 0)
        if v:
            a[3][hdr[c]] += v
    tot_i += n; tot_s += s
print(f"kernel {blk['name'][:90]}  instructions {tot_i}  samples {tot_s}")
src_cache = {}
def src(f, l):
    for d in ("rootfinding_b200/csrc",):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][l - 1].strip()[:90] if 0 < l <= len(src_cache[p]) else ""
    return ""
print("%-18s %7s %6s %6s  %-40s %s" % ("line", "inst%", "lanes", "smp%", "top stalls", "source"))
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:45]:
    top = ", ".join(f"{k[6:]}:{v}" for k, v in a[3].most_common(3))
    print("%-18s %6.2f%% %6.1f %5.2f%%  %-40s %s" % (f"{f}:{l}", 100.0 * a[0] / max(tot_i, 1), a[1] / max(a[0], 1), 100.0 * a[2] / max(tot_s, 1), top, src(f, l)))
