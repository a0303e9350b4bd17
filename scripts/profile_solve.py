"""One device solve of a batch of random 2-D degree-50 systems (for ncu / quick timing)."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
warnings.simplefilter("ignore")
from rootfinding_b200.runtime import get_engine
from rootfinding_b200.workloads import rand_coeffs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
deg = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 2
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
fma = int(os.environ.get("YR_FMA", "0"))
np.random.seed(2)
arr = rand_coeffs(dim, deg, B)
systems = [[np.ascontiguousarray(m) for m in arr[s]] for s in range(B)]
errs = [np.full(dim, 2.0 ** -52)] * B
eng = get_engine()
for rep in range(reps):
    sess = eng.session(systems, errs, use_fma=fma)
    sess.time_kernels = True
    torch.cuda.synchronize(); t = time.time()
    sess.run_jobs(sess.unit_jobs())
    torch.cuda.synchronize(); dt = time.time() - t
    st = sess.stats
    kms = sum(k[0] for k in sess.kernel_ms)
    print(f"rep {rep}: B={B} deg={deg} dim={dim} boxes {st.boxes} time {dt*1e3:.1f} ms kernels {kms:.1f} ms -> {st.boxes/dt:.3e} boxes/s "
          f"({16*st.entry_coeffs/kms/1e6:.1f} GB/s algorithmic) retries {st.retries} "
          f"rounds {getattr(st, 'rounds', 0)} launches {st.launches} scalar {getattr(sess, 'scalar_ms', 0.0):.1f} ms "
          f"zooms {st.zooms} subdiv {st.subdivisions} discards c/q/z {st.discard_const}/{st.discard_quad}/{st.discard_zoom}")
free, total = torch.cuda.mem_get_info()
print(f"device memory in use after the solves (library pools keep their high-water mark): {(total - free) / 2**30:.2f} GiB "
      f"= {(total - free) / max(B, 1) / 2**20:.1f} MiB per system; input coefficients {sum(m.size for Ms in systems for m in Ms) * 8 / 2**20 / B:.3f} MiB per system")
for (lvl, k) in zip(st.per_level, sess.kernel_ms):
    print("  level: in %7d out %7d res %6d smem %s ctas %5d thr %3d ws %6d warp %s | %8.3f ms %8.1f GB/s %9.0f boxes/ms" %
          (lvl + (k[0], k[1] / k[0] / 1e6, k[2] / k[0])))
