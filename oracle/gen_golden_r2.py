"""ORACLE -- test infrastructure only.

Round-2 fixtures: runs the UNMODIFIED reference (/root/reference, `np.product = np.prod` shim) on the inputs of
tests/cases_r2.py and writes tests/golden/solves_r2.npz:

  rnd_*      seeded random solves the round-1 fixture lacks: 2-D degree 75 / 100, 3-D degree 8 / 10, 4-D degree 4 / 5,
             5-D degree 2 (roots + solvePolyRecursive counts)
  alld_*     dimension-4 / 5 solves with all_dim_quadratic_check=True
  dev_*      the devastating example chebpolyqeps / chebspolyqeps, dim 2-3, eps 1e-1 / 1e-3 / 1e-6 (BASELINE config 5)
  nb_*       CombinedNotebook demos 1, 2, 4, the 3-D and the 5-D demo through yr.solve (roots + bounding boxes)
  ex*        4-D examples of tests/high_dim/4d_examples.py through yr.solve

    python oracle/gen_golden_r2.py [only-prefix ...]       # from the repo root, in the build container
"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
GOLD = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, HERE)

from gen_golden import import_reference, rand_coeffs, BoxCounter   # noqa: E402
import cases_r2 as cz                                               # noqa: E402


def main():
    only = sys.argv[1:]
    want = lambda tag: (not only) or any(tag.startswith(o) for o in only)
    os.chdir("/tmp")
    warnings.simplefilter("ignore")
    yroots, CSS, CA, QC = import_reference()
    counter = BoxCounter(CSS)
    path = os.path.join(GOLD, "solves_r2.npz")
    out = dict(np.load(path)) if os.path.exists(path) else {}

    def save():
        np.savez_compressed(path, **out)

    if want("rnd"):
        n = 0
        for dim, deg, N, kind in cz.RANDOM_PLAN:
            np.random.seed(2)
            batch = rand_coeffs(dim, deg, N, kind)
            for s in range(N):
                Ms = [np.ascontiguousarray(m) for m in batch[s]]
                counter.n = 0
                t0 = time.time()
                roots = CSS.solveChebyshevSubdivision([m.copy() for m in Ms], np.full(dim, 2. ** -52))
                out[f"rnd_meta_{n}"] = np.array([dim, deg, s, float(kind == "randint"), counter.n])
                out[f"rnd_roots_{n}"] = np.asarray(roots, dtype=float).reshape(-1, dim)
                print("  rnd", dim, deg, s, kind, "boxes", counter.n, "roots", len(out[f"rnd_roots_{n}"]), f"{time.time() - t0:.1f}s", flush=True)
                n += 1
        out["rnd_n"] = np.array(n)
        save()

    if want("alld"):
        n = 0
        for dim, deg, N, kind in cz.ALLDIM_PLAN:
            np.random.seed(2)
            batch = rand_coeffs(dim, deg, N, kind)
            for s in range(N):
                Ms = [np.ascontiguousarray(m) for m in batch[s]]
                counter.n = 0
                roots = CSS.solveChebyshevSubdivision([m.copy() for m in Ms], np.full(dim, 2. ** -52), all_dim_quadratic_check=True)
                nb_on = counter.n
                counter.n = 0
                CSS.solveChebyshevSubdivision([m.copy() for m in Ms], np.full(dim, 2. ** -52))
                out[f"alld_meta_{n}"] = np.array([dim, deg, s, nb_on, counter.n])
                out[f"alld_roots_{n}"] = np.asarray(roots, dtype=float).reshape(-1, dim)
                print("  alld", dim, deg, s, "boxes with/without", nb_on, counter.n, flush=True)
                n += 1
        out["alld_n"] = np.array(n)
        save()

    if want("dev"):
        for kind, dim, eps in cz.DEVASTATING:
            Ms = cz.devastating_coeffs(kind, dim, eps)
            polys = [yroots.MultiCheb(m) for m in Ms]
            counter.n = 0
            roots, boxes = yroots.solve(polys, -np.ones(dim), np.ones(dim), returnBoundingBoxes=True)
            key = cz.dev_key(kind, dim, eps)
            out[key + "_roots"] = np.asarray(roots, dtype=float).reshape(-1, dim)
            out[key + "_boxes"] = np.asarray(boxes, dtype=float).reshape(-1, dim, 2)
            out[key + "_nbox"] = np.array(counter.n)
            print("  ", key, "roots", len(out[key + "_roots"]), "boxes", counter.n, flush=True)
        save()

    for name, (funcs, a, b) in cz.NOTEBOOK.items():
        if not want("nb_" + name):
            continue
        counter.n = 0
        t0 = time.time()
        roots, boxes = yroots.solve(funcs, a, b, returnBoundingBoxes=True)
        dim = len(a)
        out[f"nb_{name}_roots"] = np.asarray(roots, dtype=float).reshape(-1, dim)
        out[f"nb_{name}_boxes"] = np.asarray(boxes, dtype=float).reshape(-1, dim, 2)
        out[f"nb_{name}_nbox"] = np.array(counter.n)
        print("  nb", name, "roots", len(out[f"nb_{name}_roots"]), "boxes", counter.n, f"{time.time() - t0:.1f}s", flush=True)
        save()

    for name, funcs in cz.FOUR_D.items():
        if not want(name):
            continue
        counter.n = 0
        t0 = time.time()
        roots = yroots.solve(funcs, -np.ones(4), np.ones(4))
        out[f"{name}_roots"] = np.asarray(roots, dtype=float).reshape(-1, 4)
        out[f"{name}_nbox"] = np.array(counter.n)
        print("  4d", name, "roots", len(out[f"{name}_roots"]), "boxes", counter.n, f"{time.time() - t0:.1f}s", flush=True)
        save()


if __name__ == "__main__":
    main()
