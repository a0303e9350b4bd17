"""Static count of local-memory instructions (LDL / STL) per source line of one kernel in a built object (-lineinfo).

    python scripts/local_mem_by_line.py OBJECT MANGLED_SUBSTRING
"""
import collections, os, re, subprocess, sys, tempfile
obj, key = os.path.abspath(sys.argv[1]), sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
cur_fn, cur_line, per = None, None, collections.Counter()
total = 0
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", ln)
    if m:
        cur_fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if cur_fn and key in cur_fn and re.search(r"\b(LDL|STL)(\.|\s)", ln):
        per[cur_line] += 1; total += 1
print(f"{key}: {total} LDL/STL instructions")
for (f, l), c in per.most_common(25):
    src = ""
    try:
        path = os.path.join(os.path.dirname(obj), f)
        src = open(path).read().splitlines()[l - 1].strip()[:110]
    except Exception:
        pass
    print(f"{c:5d}  {f}:{l}  {src}")
