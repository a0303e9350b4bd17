"""TEST INFRASTRUCTURE (fixture generator, runs only in the build container).

Runs the three alternative trim policies of the UNMODIFIED reference -- trimMsOptimized1/2/3 in
/root/reference/tests/TrimMs optimization/ChebyshevSubdivisionSolverWithOptimizedTrimMs.py:1167-1271 -- and the reference's
trimMs (:1125 of that file, identical to yroots/ChebyshevSubdivisionSolver.py:1131) on seeded tensors with decaying
coefficients, and stores inputs and outputs in tests/golden/trim_policies.npz:

    M_<case>_<p>, err_<case>            input tensors of the system and its error bounds
    shape_<case>_<policy>, err_<case>_<policy>   extents of every tensor after the trim [npoly, dim], errors after

    python oracle/gen_golden_trim.py
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
SRC = "/root/reference/tests/TrimMs optimization/ChebyshevSubdivisionSolverWithOptimizedTrimMs.py"


def main():
    np.product = np.prod                                       # numpy 2 dropped the alias the reference still uses
    sys.path.insert(0, "/root/reference")
    spec = importlib.util.spec_from_file_location("ref_trim_variants", SRC)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    funcs = {0: mod.trimMs, 1: mod.trimMsOptimized1, 2: mod.trimMsOptimized2, 3: mod.trimMsOptimized3}
    out = {}
    cases = [(2, 12, 0), (2, 20, 1), (2, 9, 2), (3, 7, 3), (3, 10, 4), (4, 5, 5), (2, 30, 6), (3, 6, 7)]
    for dim, deg, seed in cases:
        rs = np.random.RandomState(100 + seed)
        name = f"d{dim}_n{deg}_s{seed}"
        Ms = []
        for p in range(dim):
            n = [deg + 1 - rs.randint(0, 3) for _ in range(dim)]
            idx = np.indices(n).astype(float)
            rate = rs.uniform(1.0, 3.0, size=dim)
            decay = np.exp(-np.tensordot(rate, idx, axes=1))    # geometric decay, different per axis
            Ms.append(rs.randn(*n) * decay)
        errs = 10.0 ** rs.uniform(-5, -1, size=dim)
        for p, M in enumerate(Ms):
            out[f"M_{name}_{p}"] = M
        out[f"err_{name}"] = errs
        for k, fn in funcs.items():
            Mk, ek = [M.copy() for M in Ms], errs.copy()
            fn(Mk, ek)
            for p in range(dim):                                # the trims only drop trailing hyperplanes
                assert np.array_equal(Mk[p], Ms[p][tuple(slice(0, e) for e in Mk[p].shape)])
            out[f"shape_{name}_{k}"] = np.array([M.shape for M in Mk], dtype=np.int64)
            out[f"err_{name}_{k}"] = ek
            print(name, "policy", k, [M.shape for M in Mk], ek)
    out["cases"] = np.array([f"d{d}_n{n}_s{s}" for d, n, s in cases])
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "trim_policies.npz"), **out)


if __name__ == "__main__":
    main()
