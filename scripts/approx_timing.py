"""chebApproximate: opaque numpy callables (sampled point by point on the host, like the reference :69) against the same
functions as torch-vectorised device_function objects (sampled on the device grid), and the CPU oracle."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
warnings.simplefilter("ignore")
import rootfinding_b200 as yr
from rootfinding_b200 import ChebyshevApproximator as A
from oracle import approximator as oapp

def pair(xp):
    return {"readme f": (lambda x, y: xp.sin(x * y) + x * xp.log(y + 3) - x ** 2 + 1 / (y - 4), [-1., -2.], [0., 1.]),
            "readme g": (lambda x, y: xp.cos(3 * x * y) + xp.exp(3 * y / (x - 2)) - x - 6, [-1., -2.], [0., 1.]),
            "oscillatory": (lambda x, y: xp.sin(30 * x) * xp.cos(25 * y) + xp.exp(x * y) - 1.1, [-1., -1.], [1., 1.]),
            "3-D": (lambda x, y, z: xp.cos(5 * x * y) + xp.sin(4 * z + x) - y * z, [-1., -1., -1.], [1., 1., 1.])}

def timed(fn, reps=3):
    fn(); best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); out = fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    return best, out

for name in pair(np):
    fn_np, a, b = pair(np)[name]; fn_t = pair(torch)[name][0]
    a, b = np.array(a), np.array(b)
    t_host, (c_h, e_h) = timed(lambda: A.chebApproximate(fn_np, a, b))
    t_dev, (c_d, e_d) = timed(lambda: A.chebApproximate(yr.device_function(fn_t), a, b))
    t_cpu, (c_o, e_o) = timed(lambda: oapp.cheb_approximate(fn_np, a, b), reps=1) if hasattr(oapp, "cheb_approximate") else (float("nan"), (c_h, e_h))
    print(f"{name:12s} shape {str(c_h.shape):14s} host-sampled {t_host*1e3:8.1f} ms | device_function {t_dev*1e3:7.1f} ms | CPU oracle {t_cpu*1e3:8.1f} ms | "
          f"max |coeff diff| device vs host sampling {np.max(np.abs(c_h - c_d)):.1e}", flush=True)
