import os, sys, time, warnings
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
warnings.simplefilter("ignore")
from rootfinding_b200.runtime import get_engine
from rootfinding_b200.workloads import rand_coeffs
B=1024
np.random.seed(2); arr = rand_coeffs(2, 50, B)
systems = [[np.ascontiguousarray(m) for m in arr[s]] for s in range(B)]
errs = [np.full(2, 2.0 ** -52)] * B
eng = get_engine()
for rep in range(3):
    sess = eng.session(systems, errs); sess.time_kernels = True
    torch.cuda.synchronize(); t=time.time()
    jobs = sess.unit_jobs()
    t1=time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sess.run_jobs(jobs); e1.record(); torch.cuda.synchronize()
    print("unit_jobs %.2f ms, run_jobs wall %.2f ms, events %.2f ms, tensor %.1f scalar %.1f" % ((t1-t)*1e3, (time.time()-t1)*1e3, e0.elapsed_time(e1), sess.kernel_ms[0][0], sess.scalar_ms))
