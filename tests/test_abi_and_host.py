"""CPU tier: the C ABI library loads and exports what include/*.h declares, the product refuses to
run without CUDA, and the kernel-source emulator reproduces the oracle through the real host code."""
import ctypes
import os
import re
import warnings

import numpy as np
import pytest

import helpers as H
from oracle import kernels as okern
from oracle import quadratic as oquad
from oracle import subdivision as osub
from rootfinding_b200 import ChebyshevSubdivisionSolver as S
from rootfinding_b200 import QuadraticCheck as Q
from rootfinding_b200 import _lib, engine as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu():
    from emul.backend import emulated_engine
    return emulated_engine(threads=8, lanes=4)


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "yroots_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(yr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_match_binding():
    assert _declared_symbols() == sorted(_lib.EXPORTS)


def test_cuda_library_exports_every_symbol():
    """No compute calls: only dlopen + dlsym + the pure-host helpers."""
    if not os.path.exists(_lib.CUDA_LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.CUDA_LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    _lib.bind(lib)
    assert lib.yr_abi_version() == 1
    assert b"sm_100a" in lib.yr_build_info()
    p = _lib.default_params(lib)
    assert (p.max_zoom_count, p.constant_check, p.low_dim_quad_check, p.all_dim_quad_check, p.exact) == (25, 1, 1, 0, 0)
    assert p.trim_rel_tol == 1e-3
    for dim in range(1, 7):
        for kind in (0, 1, 2):
            assert lib.yr_record_layout(dim, kind, None, 0) > 0
        rd = _lib.result_dtype(lib, dim)
        assert rd["root"].shape == (dim,) and rd["box"].shape == (dim, 2)
    assert lib.yr_workspace_doubles(2, 5202, 2601, 51, 0) > 5202


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        E.CudaMemory()
    with pytest.raises(RuntimeError):
        E.SubdivisionEngine()


@pytest.mark.parametrize("dim,deg,N,kind,exact", [
    (1, 10, 3, "randn", False), (2, 3, 4, "randn", False), (2, 5, 4, "randint", False), (2, 10, 3, "randn", False),
    (2, 10, 2, "randn", True), (3, 3, 2, "randn", False), (3, 6, 1, "randn", False), (4, 2, 2, "randn", False),
    (2, 20, 2, "randn", False)])
def test_emulated_solver_matches_oracle(emu, dim, deg, N, kind, exact):
    systems = H.seeded_systems(dim, deg, N, kind)
    errs = [np.full(dim, 2. ** -52)] * N
    sess, boxes, worst = H.check_batch_against_oracle(S.solve_batch, systems, errs, engine=emu, exact=exact)
    assert sess.stats.boxes == boxes          # same search tree as the reference, box for box
    assert worst < 1e-14


def test_emulated_solver_golden_roots(emu):
    """Against roots the unmodified reference produced (tests/golden/solves_random.npz)."""
    z = H.load("solves_random.npz")
    cache = {}
    for i in range(int(z["n"])):
        dim, deg, s, randint, exact, nbox = (int(v) for v in z[f"meta_{i}"])
        if deg > 10 or dim > 3:
            continue
        key = (dim, deg, randint)
        if key not in cache:
            N = {(1, 10): 3, (2, 3): 4, (2, 5): 4, (2, 10): 3, (3, 3): 2, (3, 6): 2}[(dim, deg)]
            cache[key] = H.seeded_systems(dim, deg, N, "randint" if randint else "randn")
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            got, sess = S.solve_batch([cache[key][s]], [np.full(dim, 2. ** -52)], exact=bool(exact), engine=emu,
                                      return_session=True)
        want = z[f"roots_{i}"]
        assert len(got[0]) == len(want)
        if len(want):
            H.match_points(want, got[0], H.ROOT_TOL)
        assert sess.stats.boxes == nbox


def test_emulated_restriction_bit_exact(emu):
    """TransformChebInPlaceND through yr_init_boxes == the reference, bit for bit (plain and exact)."""
    z = H.load("restrict_axis.npz")
    groups = {}
    for i in range(int(z["n"])):
        axis, al, be, exact = z[f"par_{i}"]
        M = z[f"in_{i}"]
        if M.size > 700:
            continue
        groups.setdefault((M.ndim, bool(exact)), []).append((i, int(axis), al, be, M))
    checked = 0
    for (D, exact), items in groups.items():
        systems, alphas, betas = [], [], []
        for (i, axis, al, be, M) in items:
            systems.append([M] + [np.zeros((1,) * D)] * (D - 1))
            a, b = np.ones(D), np.zeros(D)
            a[axis], b[axis] = al, be
            alphas.append(a)
            betas.append(b)
        out, errs = S.restrict_systems(systems, [np.zeros(D)] * len(items), alphas, betas, exact=exact, engine=emu)
        for (i, axis, al, be, M), Ms in zip(items, out):
            want = z[f"out_{i}"]
            assert Ms[0].shape == want.shape, (i, Ms[0].shape, want.shape)
            assert np.array_equal(Ms[0], want), i
            checked += 1
    assert checked > 100


def test_emulated_restrict_system_errors(emu):
    """transformChebToInterval: tensors bit-exact, error bound to rounding of the |M| sums."""
    rng = np.random.RandomState(5)
    for D, shp in [(2, (9, 7)), (3, (5, 6, 4))]:
        Ms = [rng.randn(*shp) * np.exp(-0.4 * np.indices(shp).sum(0)) for _ in range(D)]
        errs = np.abs(rng.randn(D)) * 1e-12
        al, be = rng.uniform(0.05, 0.6, D), rng.uniform(-0.3, 0.3, D)
        want_M, want_E = osub.restrict_system([m.copy() for m in Ms], al, be, errs.copy(), False)
        got_M, got_E = S.transformChebToInterval(Ms, al, be, errs, False) if False else S.restrict_systems(
            [Ms], [errs], [al], [be], engine=emu)
        for a, b in zip(want_M, got_M[0]):
            assert np.array_equal(np.ascontiguousarray(a), b)
        assert np.allclose(want_E, got_E[0], rtol=1e-12, atol=0)


def test_emulated_bils_matches_reference(emu):
    z = H.load("bils.npz")
    by_dim = {}
    for i in range(int(z["n"])):
        Ms = H.unpack(z, f"Ms_{i}")
        by_dim.setdefault(len(Ms), []).append((i, Ms))
    for D, items in by_dim.items():
        iv, fl = S.bils_batched([Ms for _, Ms in items], [z[f"errs_{i}"] for i, _ in items],
                                [int(z[f"final_{i}"]) for i, _ in items], engine=emu)
        for k, (i, Ms) in enumerate(items):
            want_iv, want_fl = z[f"iv_{i}"], z[f"flags_{i}"]
            assert list(fl[k]) == list(want_fl), (i, fl[k], want_fl)
            if np.all(np.isfinite(want_iv)):
                # different SVD (one-sided Jacobi vs LAPACK) and summation order: a few ulp
                assert np.allclose(iv[k], want_iv, rtol=1e-11, atol=1e-13), (i, iv[k], want_iv)
            else:
                assert np.array_equal(np.isnan(iv[k]), np.isnan(want_iv))


def test_emulated_quadratic_check_matches_reference(emu):
    z = H.load("quadcheck.npz")
    by_dim = {}
    for i in range(int(z["n"])):
        by_dim.setdefault(z[f"M_{i}"].ndim, []).append(i)
    for D, idx in by_dim.items():
        Ms, tols = [z[f"M_{i}"] for i in idx], [float(z[f"tol_{i}"]) for i in idx]
        got = Q.quadratic_check_batched(Ms, tols, engine=emu)
        got_nd = Q.quadratic_check_batched(Ms, tols, nd_check=True, engine=emu)
        want = np.array([z[f"res_{i}"] for i in idx])
        assert np.array_equal(got, want[:, 0]), D
        assert np.array_equal(got_nd, want[:, 1]), D


def test_engine_overflow_retry_and_global_workspace(emu):
    """Force the optimistic-capacity retry and the global-memory workspace path."""
    systems = H.seeded_systems(2, 10, 2)
    errs = [np.full(2, 2. ** -52)] * 2
    old = (emu.worst_case_budget, emu.max_smem_cta)
    pipe = emu.pipeline
    try:
        emu.pipeline = 1                        # the level-synchronous pipeline (serves dim != 2 and exact mode)
        emu.worst_case_budget = 0               # never allocate the worst case up front
        emu.max_smem_cta = 1024                 # nothing fits "shared memory"
        sess, boxes, worst = H.check_batch_against_oracle(S.solve_batch, systems, errs, engine=emu)
        assert sess.stats.boxes == boxes
        assert not any(lvl[3] for lvl in sess.stats.per_level)       # every level ran from the global workspace
    finally:
        emu.worst_case_budget, emu.max_smem_cta = old
        emu.pipeline = pipe


def test_argument_errors():
    with pytest.raises(ValueError, match="Solver Takes in N polynomials of dimension N!"):
        S.solveChebyshevSubdivision([np.zeros((2, 2)), np.zeros(3)], np.zeros(2))
    with pytest.raises(ValueError, match="Ms and errors must be same length!"):
        S.solveChebyshevSubdivision([np.zeros((2, 2)), np.zeros((2, 2))], np.zeros(3))


@pytest.mark.parametrize("dim,deg,N", [(1, 10, 2), (2, 10, 3), (3, 4, 2), (4, 2, 1)])
def test_emulated_warp_per_box(emu, dim, deg, N):
    """The level pipeline's warp-per-box kernels give the same search tree and roots."""
    systems = H.seeded_systems(dim, deg, N)
    errs = [np.full(dim, 2. ** -52)] * N
    old = emu.warp_min_items
    pipe = emu.pipeline
    try:
        emu.pipeline = 1
        emu.warp_min_items = 1                  # every level runs warp-per-box
        sess, boxes, worst = H.check_batch_against_oracle(S.solve_batch, systems, errs, engine=emu)
        assert sess.stats.boxes == boxes and worst < 1e-14
        assert all(lvl[7] for lvl in sess.stats.per_level)
    finally:
        emu.warp_min_items = old
        emu.pipeline = pipe


def test_emulated_warp_shuffle_matrix_build(emu):
    """32-lane warp-teams exercise the shuffle-based restriction-matrix build (extents <= warp width);
    the search tree and roots must not change, and restrictions stay bit-exact."""
    old = (emu.warp_min_items, os.environ.get("YR_EMUL_LANES"), os.environ.get("YR_EMUL_THREADS"))
    pipe = emu.pipeline
    try:
        emu.pipeline = 1
        emu.warp_min_items = 1
        os.environ["YR_EMUL_LANES"] = "32"
        os.environ["YR_EMUL_THREADS"] = "32"
        systems = H.seeded_systems(2, 7, 2)
        errs = [np.full(2, 2. ** -52)] * 2
        sess, boxes, worst = H.check_batch_against_oracle(S.solve_batch, systems, errs, engine=emu)
        assert sess.stats.boxes == boxes and worst < 1e-14
        assert all(lvl[7] for lvl in sess.stats.per_level)
    finally:
        emu.warp_min_items = old[0]
        emu.pipeline = pipe
        os.environ["YR_EMUL_LANES"] = old[1] or "4"
        os.environ["YR_EMUL_THREADS"] = old[2] or "8"


@pytest.mark.parametrize("deg,N,kind", [(5, 4, "randn"), (10, 3, "randn"), (12, 2, "randint"), (20, 2, "randn")])
def test_emulated_frontier_pipeline(emu, deg, N, kind):
    """The round-based 2-D frontier pipeline (yr_frontier_solve): same search tree and roots as the oracle."""
    systems = H.seeded_systems(2, deg, N, kind)
    errs = [np.full(2, 2. ** -52)] * N
    assert emu.pipeline == 2
    sess, boxes, worst = H.check_batch_against_oracle(S.solve_batch, systems, errs, engine=emu)
    assert sess.use_frontier and sess.stats.rounds > 0
    assert sess.stats.boxes == boxes and worst < 1e-14


def test_emulated_frontier_matches_level_pipeline(emu):
    """Both device pipelines visit the same boxes and report identical records (restrictions are bit-exact in both)."""
    systems = H.seeded_systems(2, 9, 3)
    errs = [np.full(2, 2. ** -52)] * 3
    pipe = emu.pipeline
    try:
        out = []
        for p in (2, 1):
            emu.pipeline = p
            roots, sess = S.solve_batch(systems, errs, engine=emu, return_session=True)
            out.append((roots, sess.stats.boxes, sess.stats.zooms, sess.stats.subdivisions))
        assert out[0][1:] == out[1][1:]
        for a, b in zip(out[0][0], out[1][0]):
            assert a.shape == b.shape and np.max(np.abs(a - b), initial=0.0) < 1e-15
    finally:
        emu.pipeline = pipe


def test_emulated_frontier_fma_and_unit_extents(emu):
    """use_fma and degenerate shapes (a polynomial that is constant along an axis) through the frontier pipeline."""
    rng = np.random.RandomState(5)
    f = rng.randn(6, 1)                          # depends on x only
    g = rng.randn(1, 7)                          # depends on y only
    errs = [np.full(2, 2. ** -52)]
    sess, boxes, worst = H.check_batch_against_oracle(S.solve_batch, [[f, g]], errs, engine=emu)
    assert sess.use_frontier and sess.stats.boxes == boxes
    systems = H.seeded_systems(2, 8, 2)
    got = S.solve_batch(systems, [np.full(2, 2. ** -52)] * 2, use_fma=True, engine=emu)
    want = S.solve_batch(systems, [np.full(2, 2. ** -52)] * 2, engine=emu)
    for a, b in zip(got, want):
        assert a.shape == b.shape and np.max(np.abs(a - b), initial=0.0) < 1e-10


def test_emulated_hybrid_level_then_frontier(emu):
    """Large 2-D systems start on the level pipeline and move to the frontier pipeline once their boxes fit its
    shared-memory budget; here the budget is shrunk so a small problem takes that route.  Records of both
    pipelines must come back in production order and the reporting order must match the oracle's."""
    class Shim:
        def __init__(self, lib):
            self.lib, self.asked = lib, 0

        def __getattr__(self, k):
            if k == "yr_frontier_fits_dim":
                def fits(dim, extent, tensor):          # "too large" for the first three levels of every run
                    self.asked += 1
                    return int(self.asked % 4 == 0)
                return fits
            return getattr(self.lib, k)
    systems = H.seeded_systems(2, 12, 3)
    errs = [np.full(2, 2. ** -52)] * 3
    real = emu.lib
    try:
        emu.lib = Shim(real)
        sess, boxes, worst = H.check_batch_against_oracle(S.solve_batch, systems, errs, engine=emu)
    finally:
        emu.lib = real
    assert sess.stats.boxes == boxes and worst < 1e-14
    assert sess.stats.rounds > 0 and sess.stats.levels > sess.stats.rounds      # both pipelines took part
    ref = S.solve_batch(systems, errs, engine=emu)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        emu.lib = Shim(real)
        try:
            got = S.solve_batch(systems, errs, engine=emu)
        finally:
            emu.lib = real
    for a, b in zip(got, ref):
        assert a.shape == b.shape and np.max(np.abs(a - b), initial=0.0) < 1e-14    # same roots in the same (depth-first) order


def test_slim_result_path_matches_full_records():
    """solve_batch answers ordinary batches from the device-gathered (system, root) entries alone; the full result
    records (requested with returnBoundingBoxes) must give the same roots in the same order."""
    from emul.backend import emulated_engine
    from rootfinding_b200 import ChebyshevSubdivisionSolver as S
    from oracle.gen_golden import rand_coeffs
    eng = emulated_engine()
    np.random.seed(5)
    arr = rand_coeffs(2, 9, 6)
    systems = [[np.ascontiguousarray(m) for m in arr[s]] for s in range(6)]
    errs = [np.full(2, 2.0 ** -52)] * 6
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        slim, s1 = S.solve_batch(systems, errs, engine=eng, return_session=True)
        (full, boxes), s2 = S.solve_batch(systems, errs, returnBoundingBoxes=True, engine=eng, return_session=True)
    assert s1._results_host.n == 0, "the slim path must not download the full records"
    assert s2._results_host.n == s2.n_results > 0
    for a, b, bx in zip(slim, full, boxes):
        assert np.array_equal(np.asarray(a), np.asarray(b)) and len(bx) == len(b)
    assert sum(len(r) for r in slim) > 10


def test_session_is_freed_without_the_cyclic_collector():
    """A finished session must die by reference counting: it borrows the engine's page-locked download buffers, and
    a live previous borrower costs the next solve a copy of everything it downloaded."""
    import gc
    import weakref
    from emul.backend import emulated_engine
    from rootfinding_b200 import ChebyshevSubdivisionSolver as S
    from oracle.gen_golden import rand_coeffs
    eng = emulated_engine()
    np.random.seed(2)
    arr = rand_coeffs(2, 8, 2)
    systems = [[arr[s][0], arr[s][1]] for s in range(2)]
    gc.disable()
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            roots, sess = S.solve_batch(systems, [np.full(2, 2.0 ** -52)] * 2, engine=eng, return_session=True)
        ref = weakref.ref(sess)
        del sess
        assert ref() is None
    finally:
        gc.enable()


def test_workload_generator_matches_reference_stream():
    """rootfinding_b200.workloads.rand_coeffs (what bench.py feeds the GPU) draws the same numbers as the generator
    the golden vectors were made with (oracle/gen_golden.py, restating tests/random_tests.py:5-19)."""
    from rootfinding_b200.workloads import rand_coeffs, random_systems
    from oracle.gen_golden import rand_coeffs as ref
    for dim, deg, N, kind in [(2, 50, 2, "randn"), (2, 7, 3, "randint"), (3, 4, 2, "randn"), (1, 9, 2, "randn")]:
        np.random.seed(2); a = rand_coeffs(dim, deg, N, kind)
        np.random.seed(2); b = ref(dim, deg, N, kind)
        assert a.shape == (N, dim) + (deg + 1,) * dim and np.array_equal(a, b)
    systems, errs = random_systems(2, 6, 3)
    assert len(systems) == 3 and systems[0][1].shape == (7, 7) and errs[0][0] == 2.0 ** -52
    assert systems[0][0][6, 1] == 0.0 and systems[0][0][5, 1] != 0.0          # zero beyond total degree


def test_npy_test_files_round_trip(tmp_path):
    """The reference's random-test harness stores systems as .npy arrays [N, dim, deg+1, ...] (tests/random_tests.py:27-44);
    workloads.save_tests / load_tests / load_systems read and write the same files."""
    from rootfinding_b200 import workloads as W, MultiCheb, MultiPower
    fmt = str(tmp_path / "dim{dim}_deg{deg}_{kind}.npy")
    np.random.seed(2)
    W.save_tests(2, [3, 5], 4, "randint", filefmt=fmt)
    np.random.seed(2)
    want = W.gen_tests(2, [3, 5], 4, "randint")
    for deg, arr in zip([3, 5], want):
        assert np.array_equal(np.load(fmt.format(dim=2, deg=deg, kind="randint")), arr)
        tests = W.load_tests(2, deg, "power", "randint", N=3, filefmt=fmt)
        assert len(tests) == 3 and all(isinstance(p, MultiPower) for t in tests for p in t)
        cheb = W.load_tests(2, deg, "cheb", "randint", filefmt=fmt)
        assert len(cheb) == 4 and isinstance(cheb[0][0], MultiCheb)
        systems, errs = W.load_systems(2, deg, "randint", filefmt=fmt)
        assert len(systems) == 4 and np.array_equal(systems[1][0], arr[1][0]) and errs[0][0] == 2.0 ** -52


def test_batch_that_does_not_fit_is_solved_in_halves(monkeypatch):
    """A device out-of-memory failure of a batch falls back to solving its halves one after the other (systems never
    interact); a single system that does not fit still raises."""
    from emul.backend import emulated_engine
    from rootfinding_b200 import ChebyshevSubdivisionSolver as S
    from rootfinding_b200.workloads import random_systems
    eng = emulated_engine()
    systems, errs = random_systems(2, 6, 5, seed=3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = S.solve_batch(systems, errs, engine=eng)
    real, calls = eng.session, []

    from rootfinding_b200._lib import DeviceOutOfMemory

    def limited(sy, er, **kw):
        calls.append(len(sy))
        if len(sy) > 2:
            raise DeviceOutOfMemory("yroots_b200: yr_frontier_solve failed: device allocation failed (frontier buffers)")
        return real(sy, er, **kw)
    monkeypatch.setattr(eng, "session", limited)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got, sess = S.solve_batch(systems, errs, engine=eng, return_session=True)
        (got_b, boxes), _ = S.solve_batch(systems, errs, returnBoundingBoxes=True, engine=eng, return_session=True)
    assert calls[:4] == [5, 2, 3, 1] and len(sess.chunk_sessions) == 3
    for a, b, c, bx in zip(want, got, got_b, boxes):
        assert np.array_equal(a, b) and np.array_equal(a, c) and len(bx) == len(a)
    monkeypatch.setattr(eng, "session", lambda sy, er, **kw: (_ for _ in ()).throw(DeviceOutOfMemory("out of device memory")))
    with pytest.raises(DeviceOutOfMemory):
        S.solve_batch(systems[:1], errs[:1], engine=eng)
    # any other failure is not mistaken for a memory problem, whatever its text says
    monkeypatch.setattr(eng, "session", lambda sy, er, **kw: (_ for _ in ()).throw(RuntimeError("kernel failed: out of memory?")))
    with pytest.raises(RuntimeError):
        S.solve_batch(systems, errs, engine=eng)


def test_batches_are_pre_split_by_the_capacity_estimate(monkeypatch):
    """solve_batch cuts a batch whose estimated peak frontier (input bytes x a measured factor per dimension) exceeds the
    free device memory into consecutive sub-batches before launching anything."""
    from emul.backend import emulated_engine
    from rootfinding_b200 import ChebyshevSubdivisionSolver as S
    from rootfinding_b200.workloads import random_systems
    eng = emulated_engine()
    systems, errs = random_systems(2, 6, 6, seed=5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = S.solve_batch(systems, errs, engine=eng)
        per_system = 2 * 49 * 8 * S._PEAK_FACTOR[2]
        monkeypatch.setattr(eng.mem, "free_bytes", lambda: 2.5 * per_system / 0.6, raising=False)
        got, sess = S.solve_batch(systems, errs, engine=eng, return_session=True)
    assert [s.nsys for s in sess.chunk_sessions] == [2, 2, 2]
    for a, b in zip(want, got):
        assert np.array_equal(a, b)


def test_very_large_batches_are_pre_split(monkeypatch):
    """Batches above the engine's coefficient budget are solved as consecutive sub-batches; same answer."""
    from emul.backend import emulated_engine
    from rootfinding_b200 import ChebyshevSubdivisionSolver as S
    from rootfinding_b200.workloads import random_systems
    eng = emulated_engine()
    systems, errs = random_systems(2, 6, 7, seed=4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = S.solve_batch(systems, errs, engine=eng)
        monkeypatch.setattr(eng, "max_batch_coeffs", 3 * 98, raising=False)        # 98 coefficients per system
        got, sess = S.solve_batch(systems, errs, engine=eng, return_session=True)
    assert len(sess.chunk_sessions) == 3 and sum(c.nsys for c in sess.chunk_sessions) == 7
    assert sum(c.stats.boxes for c in sess.chunk_sessions) > 0
    for a, b in zip(want, got):
        assert np.array_equal(a, b)


def test_host_pool_reuses_a_buffer_only_when_nothing_refers_to_it():
    """engine.HostPool hands result arrays to callers: a buffer may come back only after every array and view made from it is
    gone (its lifetime is tracked through the arrays' base object)."""
    import gc
    from rootfinding_b200.engine import HostPool
    pool = HostPool()
    a = pool.take(4000).view(np.float64).reshape(-1, 5)
    a[...] = 7.0
    parts = [a[:30], a[30:]]
    addr = a.ctypes.data
    del a
    gc.collect()
    assert pool.free == []                                   # the slices still refer to it
    b = pool.take(4000)
    assert b.ctypes.data != addr and np.all(parts[0] == 7.0)
    del parts
    gc.collect()
    assert len(pool.free) == 1
    c = pool.take(1000)
    assert c.ctypes.data == addr and pool.free == []


def test_chunked_solve_releases_each_chunk(emu, monkeypatch):
    """a chunked solve must not keep every chunk's device buffers until the end."""
    from rootfinding_b200 import ChebyshevSubdivisionSolver as S
    from rootfinding_b200.engine import SolveSession
    systems, errs = H.seeded_systems(2, 5, 6), [np.full(2, 2.0 ** -52)] * 6
    released = []
    orig = SolveSession.release_device

    def spy(self):
        orig(self)
        released.append((self.d_coeffs, self.d_results, len(self._pending)))
    monkeypatch.setattr(SolveSession, "release_device", spy)
    whole = S.solve_batch(systems, errs, engine=emu)
    parts = S._solve_chunks(systems, errs, [(0, 2), (2, 6)], False, False, True, True, False, False, emu, False)
    assert released == [(None, None, 0), (None, None, 0)]
    for a, b in zip(whole, parts):
        assert np.array_equal(a, b)
    released.clear()
    out, last = S._solve_chunks(systems, errs, [(0, 2), (2, 6)], False, False, True, True, False, False, emu, True)
    assert released == [] and len(last.chunk_sessions) == 2 and last.chunk_sessions[0].d_coeffs is not None
