"""Host driver of the device-resident subdivision frontier.

A solve is a sequence of level launches (include/yroots_b200.h: yr_init_boxes, then
yr_process_level until the frontier is empty).  Boxes, their coefficient tensors, the
result records and the node table live in device memory for the whole solve; the host
only reads one small counter block per level to size the next launch.  PyTorch supplies
device memory (caching allocator), the stream and pinned staging buffers -- plumbing; all
arithmetic happens in the hand-written CUDA kernels behind the C ABI.

Reference: this replaces the recursion of
/root/reference/yroots/ChebyshevSubdivisionSolver.py:1177-1317 (solvePolyRecursive) and the
entry point :1387-1464; the tree-structured bookkeeping that the recursion does on its way
back up (:1266-1298 final-step de-duplication, :1318-1385 merging of touching exterior boxes)
is rootfinding_b200/postpass.py.
"""
import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import _lib

FIRST_SPLIT = 0.0394555475981047      # TrackedInterval.nextTransformPoints, ChebyshevSubdivisionSolver.py:414


class _Buf:
    __slots__ = ("ptr", "nbytes", "owner")

    def __init__(self, ptr, nbytes, owner):
        self.ptr, self.nbytes, self.owner = ptr, nbytes, owner


class _Lease:
    """Exports a pooled buffer to numpy (`np.asarray(lease)`): every array and view made from it keeps the lease alive, and
    the buffer goes back to the pool when the last of them is gone."""

    def __init__(self, storage):
        self.storage = storage

    @property
    def __array_interface__(self):
        return self.storage.__array_interface__


class HostPool:
    """Host buffers for results handed to the caller.  A fresh 90 MB array costs ~20 ms of first-touch page faults per
    call; a buffer whose previous result arrays have all been dropped is reused instead.  `alloc(nbytes)` returns the
    uint8 storage array (default: pageable numpy memory; distributed.py passes a page-locked allocator so that device ->
    host copies land in the result buffer directly)."""

    def __init__(self, keep=3, alloc=None):
        self.free, self.keep = [], keep
        self.alloc = alloc or (lambda nbytes: np.empty(nbytes, dtype=np.uint8))

    def _give_back(self, storage):
        if len(self.free) < self.keep:
            self.free.append(storage)

    def take(self, nbytes):
        import weakref
        for i, st in enumerate(self.free):
            if st.nbytes >= nbytes:
                self.free.pop(i)
                break
        else:
            st = self.alloc(nbytes + nbytes // 4)
        lease = _Lease(st)
        weakref.finalize(lease, self._give_back, st)
        return np.asarray(lease)[:nbytes]


result_pool = HostPool()


class _SparsePool:
    """Keeps the big mostly-zero tables of SolveSession.sparse_records between solves: the first-touch page faults of a
    fresh mapping cost more than the records written into it (15 000 records scattered over 140 MB: 15 ms, with or
    without transparent huge pages).  A table is lent to one session at a time; what that session wrote is zeroed
    again when the table is lent out next."""

    def __init__(self):
        self.slots = {}

    def borrow(self, slot, count, dtype, who):
        import weakref
        ent = self.slots.get(slot)
        if ent is not None:
            prev = ent["owner"]() if ent["owner"] is not None else None
            if prev is not None and prev is not who:
                prev._detach_host(slot)                            # still in use: its records move to private memory
            lease = ent["lease"]
            arr = ent["arr"]
            if arr.dtype != dtype or len(arr) < count:
                ent = None
            else:
                rows = arr.view(np.uint8).reshape(len(arr), dtype.itemsize)      # (byte rows: far cheaper than record copies)
                for idx in lease["idx"]:
                    rows[idx] = 0
                rows[lease["lo"]:lease["hi"]] = 0
        if ent is None:
            ent = self.slots[slot] = dict(arr=np.zeros(count + count // 2, dtype=dtype))
        ent["owner"] = weakref.ref(who)
        ent["lease"] = dict(idx=[], lo=0, hi=0)
        return ent["arr"], ent["lease"]


class _Growable:
    """Append-only structured host array on top of a byte buffer obtained from `alloc(nbytes, who)`."""

    def __init__(self, dtype, alloc, who):
        import weakref
        # weak: the owner holds this object; a cycle would keep a finished session (and the page-locked buffers it
        # borrowed) alive until the cyclic collector runs
        self.dtype, self.alloc, self.who = dtype, alloc, weakref.ref(who)
        self.buf = np.empty(0, dtype=np.uint8)
        self.n = 0

    def tail(self, k):
        """Room for k more records: returns the uint8 view to fill."""
        item = self.dtype.itemsize
        need = (self.n + k) * item
        if need > self.buf.nbytes:
            new = self.alloc(max(need + need // 8, 2 * self.buf.nbytes), self.who())
            new[:self.n * item] = self.buf[:self.n * item]
            self.buf = new
            self.lease = None                                      # (a lent table was left behind: the pool zeroes what it knows)
        v = self.buf[self.n * item:need]
        self.n += k
        if getattr(self, "lease", None) is not None:
            self.lease["hi"] = self.n
        return v

    def view(self):
        return self.buf[:self.n * self.dtype.itemsize].view(self.dtype)

    def adopt(self, table, n, lease=None):
        """Replace the contents by the first n records of `table` (a structured array with spare room behind them).
        `lease`: the pool's note of what this borrower writes (_SparsePool)."""
        self.buf = table.view(np.uint8).reshape(-1)
        self.n = n
        self.lease = lease
        if lease is not None:
            lease["lo"] = lease["hi"] = n

    def detach(self):
        """Move to private memory (the lender wants its buffer back)."""
        self.buf = self.buf[:self.n * self.dtype.itemsize].copy()
        self.lease = None

    def __len__(self):
        return self.n


class CudaMemory:
    """Device memory, stream and host staging through torch (cuda only)."""

    def __init__(self, device=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("yroots_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._staging = []
        self._pinned = {}
        self._pinned_owner = {}

    def empty(self, nbytes):
        n8 = max(1, (int(nbytes) + 7) // 8)
        t = self.torch.empty(n8, dtype=self.torch.int64, device=self.device)
        return _Buf(t.data_ptr(), n8 * 8, t)

    def zeros(self, nbytes):
        b = self.empty(nbytes)
        b.owner.zero_()
        return b

    def zero_(self, buf):
        buf.owner.zero_()

    def upload(self, arr):
        """Host -> device through a pinned staging buffer (asynchronous on the current stream)."""
        arr = np.ascontiguousarray(arr)
        b = self.empty(arr.nbytes)
        if arr.nbytes:
            pinned = self.torch.empty(arr.nbytes, dtype=self.torch.uint8, pin_memory=True)
            pinned.numpy()[:] = arr.view(np.uint8).reshape(-1)
            b.owner.view(self.torch.uint8)[:arr.nbytes].copy_(pinned, non_blocking=True)
            self._staging.append(pinned)                       # keep alive until the next synchronize
            self.h2d_bytes += arr.nbytes
        return b

    def upload_concat(self, chunks, total):
        """Host -> device of the concatenation of 1-D float64 `chunks` (`total` elements): gathered straight into the
        pinned staging buffer, so the packed tensor data is written once on the host."""
        b = self.empty(total * 8)
        if total:
            # one page-locked staging buffer is kept (grow-only): allocating 40 MB of pinned memory per call costs
            # several milliseconds.  The event says when the previous copy out of it has finished.
            stage = getattr(self, "_concat_stage", None)
            if stage is None or stage[0].numel() < total * 8:
                stage = self._concat_stage = (self.torch.empty(total * 8 + total, dtype=self.torch.uint8, pin_memory=True),
                                              self.torch.cuda.Event())
            else:
                stage[1].synchronize()
            pinned = stage[0][:total * 8]
            dst = pinned.numpy().view(np.float64)
            if total >= (1 << 21) and len(chunks) >= 64:
                # a 40 MB gather runs at one core's copy rate (6 ms): four threads (numpy copies release the GIL)
                pool = getattr(self, "_copy_pool", None)
                if pool is None:
                    from concurrent.futures import ThreadPoolExecutor
                    pool = self._copy_pool = ThreadPoolExecutor(4)
                ends = np.cumsum([c.size for c in chunks])
                cuts = [0] + [int(np.searchsorted(ends, total * q / 4)) + 1 for q in (1, 2, 3)] + [len(chunks)]
                jobs = []
                for g0, g1 in zip(cuts[:-1], cuts[1:]):
                    if g1 > g0:
                        lo = int(ends[g0 - 1]) if g0 else 0
                        jobs.append(pool.submit(np.concatenate, chunks[g0:g1], out=dst[lo:int(ends[g1 - 1])]))
                for j in jobs:
                    j.result()
            else:
                np.concatenate(chunks, out=dst)
            b.owner.view(self.torch.uint8)[:total * 8].copy_(pinned, non_blocking=True)
            stage[1].record(self.torch.cuda.current_stream(self.device))
            self.h2d_bytes += total * 8
        return b

    def host_bytes(self, nbytes, who=None, slot=0):
        """Host buffer for record downloads.  One page-locked buffer per slot (0 results, 1 nodes) is kept per engine
        and lent to the session that downloads (a 200 MB device -> pageable copy costs several times the solve); a
        previous borrower that is still alive first moves its records to private memory."""
        import weakref
        if nbytes < (1 << 20):
            return np.empty(nbytes, dtype=np.uint8)
        owner = self._pinned_owner.get(slot)
        prev = owner() if owner is not None else None
        if prev is not None and prev is not who:
            prev._detach_host(slot)
        buf = self._pinned.get(slot)
        if buf is None or buf.numel() < nbytes:
            self._pinned[slot] = None
            buf = self._pinned[slot] = self.torch.empty(int(nbytes), dtype=self.torch.uint8, pin_memory=True)
        self._pinned_owner[slot] = weakref.ref(who) if who is not None else None
        return buf.numpy()[:nbytes]

    def timer(self):
        """CUDA event pair helper: returns (record_start, record_stop, elapsed_ms) bound to the current stream."""
        ev = self.torch.cuda.Event
        return ev(enable_timing=True), ev(enable_timing=True)

    def download(self, buf, nbytes, offset=0):
        if nbytes == 0:
            return np.empty(0, dtype=np.uint8)
        t = buf.owner.view(self.torch.uint8)[offset:offset + nbytes]
        self.d2h_bytes += nbytes
        out = t.cpu().numpy()                                  # synchronises the stream
        self._staging.clear()
        return out

    def copy(self, dst, src, nbytes):
        dst.owner.view(self.torch.uint8)[:nbytes].copy_(src.owner.view(self.torch.uint8)[:nbytes])

    def stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def free_bytes(self):
        """Device memory a new solve can count on: everything this process does not hold in live torch tensors.  (The
        library's own grow-only pools are reused by the next solve, so they count as available.)  cudaMemGetInfo costs
        3 - 40 ms per call on a busy device, so what OTHER owners hold (context, other processes) is measured once."""
        if getattr(self, "_foreign_bytes", None) is None:
            free, total = self.torch.cuda.mem_get_info(self.device)
            self._total_bytes = total
            self._foreign_bytes = max(0, total - free - self.torch.cuda.memory_reserved(self.device))
        return self._total_bytes - self._foreign_bytes - self.torch.cuda.memory_allocated(self.device)

    def synchronize(self):
        self.torch.cuda.current_stream(self.device).synchronize()
        self._staging.clear()


@dataclass
class SolveStats:
    boxes: int = 0
    zooms: int = 0
    subdivisions: int = 0
    discard_const: int = 0
    discard_quad: int = 0
    discard_zoom: int = 0
    entry_coeffs: int = 0
    restrict_macs: int = 0
    max_level: int = 0
    levels: int = 0
    launches: int = 0
    retries: int = 0
    per_level: list = field(default_factory=list)     # (boxes in, children out, results, smem?, ctas, threads)

    def add(self, c):
        for k in ("boxes", "zooms", "subdivisions", "discard_const", "discard_quad", "discard_zoom",
                  "entry_coeffs", "restrict_macs"):
            setattr(self, k, getattr(self, k) + int(c[k]))
        self.max_level = max(self.max_level, int(c["max_level"]))

    @property
    def algorithmic_bytes(self):
        """every box's tensors are written once and read once."""
        return 16 * self.entry_coeffs


class SubdivisionEngine:
    """Owns the library handle and launch heuristics; create sessions to solve batches."""

    def __init__(self, lib=None, mem=None, threads=None, max_ctas_per_sm=None):
        self.lib = _lib.load_cuda_library() if lib is None else lib
        self.mem = CudaMemory() if mem is None else mem
        sm, smem_cta, smem_sm = C.c_int32(), C.c_int64(), C.c_int64()
        _lib.check(self.lib, self.lib.yr_device_limits(C.byref(sm), C.byref(smem_cta), C.byref(smem_sm)), "yr_device_limits")
        self.sm_count, self.max_smem_cta, self.smem_per_sm = sm.value, smem_cta.value, smem_sm.value
        self.threads = threads
        self.max_ctas_per_sm = max_ctas_per_sm
        self.worst_case_budget = 1 << 28              # below this many bytes, just allocate the worst case
        self.warp_max_ws = int(os.environ.get("YR_WARP_MAX_WS", "28000"))     # doubles; larger boxes run CTA-per-box
        self.warp_min_items = int(os.environ.get("YR_WARP_MIN_ITEMS", "1024"))  # fewer boxes than this: CTA-per-box (latency)
        self.warp_teams = int(os.environ.get("YR_WARP_TEAMS", "16"))          # warp-teams per SM (register bound: 16 x 32 x 128)
        # 2: the round-based frontier pipelines (yr_frontier_solve: dimension 2, 3, 4, plain arithmetic) where they apply,
        #    the level pipeline elsewhere; 1: the level pipeline (phase-split kernels per level) for everything
        self.pipeline = max(1, int(os.environ.get("YR_PIPELINE", "2")))
        self.sparse_pool = _SparsePool()

    def params(self, **kw):
        p = _lib.default_params(self.lib)
        for k, v in kw.items():
            if not hasattr(p, k):
                raise TypeError(f"unknown solver parameter {k}")
            setattr(p, k, type(getattr(p, k))(v))
        return p

    def launch_config(self, dim, n_items, total, max_tensor, max_extent, exact, allow_warp=True):
        """threads / persistent CTAs / workspace placement / box granularity for boxes of at most this size.

        Small boxes run warp-per-box (every warp of a 128-thread CTA owns its own box: no block barriers,
        a serial section stalls one warp only); larger ones CTA-per-box; boxes that do not fit shared
        memory run CTA-per-box from a global workspace.
        """
        ws = int(self.lib.yr_workspace_doubles(dim, int(total), int(max_tensor), int(max_extent), int(exact)))
        static = int(self.lib.yr_static_smem_bytes(dim))
        per_warp = static + 16 + ((ws * 8 + 15) & ~15)
        if allow_warp and ws <= self.warp_max_ws and n_items >= self.warp_min_items and per_warp + 1024 <= self.max_smem_cta:
            # independent warp-teams, 4 per CTA (fewer when a box is large)
            warps = int(max(1, min(4, (self.max_smem_cta - 1024) // per_warp)))
            per_cta = warps * per_warp + 1024
            per_sm = max(1, min(self.warp_teams // warps, self.smem_per_sm // per_cta))
            ctas = int(max(1, min((n_items + warps - 1) // warps, self.sm_count * per_sm)))
            return 32 * warps, ctas, True, ws, 1
        threads = self.threads or (128 if total >= 768 else 64)
        in_smem = ws * 8 + static <= self.max_smem_cta
        if in_smem:
            per_sm = max(1, min(32, 2048 // threads, self.smem_per_sm // (ws * 8 + static + 1024)))
        else:
            per_sm = 2
        if self.max_ctas_per_sm:
            per_sm = min(per_sm, self.max_ctas_per_sm)
        ctas = int(max(1, min(n_items, self.sm_count * per_sm)))
        return threads, ctas, in_smem, ws, False

    def release_cached(self):
        """Give cached device memory back (after an out-of-memory failure, before retrying with a smaller batch)."""
        try:
            self.lib.yr_frontier_trim(self.mem.stream())
        except Exception:
            pass
        t = getattr(self.mem, "torch", None)
        if t is not None:
            t.cuda.synchronize()
            t.cuda.empty_cache()

    def session(self, systems, errors, **params):
        return SolveSession(self, systems, errors, self.params(**params))


class SolveSession:
    """One batch of systems resident on the device, plus everything solves on it produce."""

    def __init__(self, engine, systems, errors, prm):
        self.eng, self.lib, self.mem, self.prm = engine, engine.lib, engine.mem, prm
        self.dim = D = len(systems[0])
        if not 1 <= D <= 6:
            raise ValueError("yroots_b200 supports systems of dimension 1..6")
        self.nsys = len(systems)
        chunks, shape_list = [], []
        f64 = np.dtype(np.float64)
        for Ms in systems:                                        # one pass: this loop is on the end-to-end path
            if len(Ms) != D:
                raise ValueError("Solver Takes in N polynomials of dimension N!")
            for M in Ms:
                if type(M) is not np.ndarray or M.dtype != f64:
                    M = np.asarray(M, dtype=np.float64)
                if M.ndim != D:
                    raise ValueError("Solver Takes in N polynomials of dimension N!")
                shape_list.append(M.shape)
                chunks.append(M.reshape(-1))                      # copies only a non-contiguous tensor
        shapes = np.array(shape_list, dtype=np.int32).reshape(self.nsys, D, D) if shape_list else np.ones((0, D, D), np.int32)
        sizes = shapes.prod(axis=2, dtype=np.int64).reshape(-1)
        offs = (np.cumsum(sizes) - sizes).reshape(self.nsys, D)
        pos = int(sizes.sum())
        self.shapes_h, self.total_coeffs = shapes, pos
        self.sys_total = shapes.prod(axis=2).sum(axis=1)          # coefficients per system
        self.sys_max_tensor = shapes.prod(axis=2).max(axis=1)
        self.sys_max_extent = shapes.reshape(self.nsys, -1).max(axis=1)
        errs = np.ascontiguousarray(np.asarray(errors, dtype=np.float64).reshape(self.nsys, D))
        m = self.mem
        if hasattr(m, "upload_concat"):
            self.d_coeffs = m.upload_concat(chunks, pos)
        else:
            self.d_coeffs = m.upload(np.concatenate(chunks) if chunks else np.zeros(0))
        self.d_off = m.upload(offs)
        self.d_shapes = m.upload(shapes)
        self.d_errs = m.upload(errs)
        self.rdt = _lib.result_dtype(self.lib, D)
        self.ndt = _lib.node_dtype(self.lib, D)
        self.box_bytes = int(self.lib.yr_record_layout(D, 0, None, 0))
        self.d_counters = m.zeros(_lib.COUNTER_BYTES)
        self.cap_results = 0
        self.n_results = 0
        self.d_results = None
        self.cap_nodes = 0
        self.n_nodes = 0
        self.d_nodes = None
        self._pending = []
        alloc = getattr(m, "host_bytes", lambda nbytes, who=None: np.empty(nbytes, dtype=np.uint8))
        self._results_host = _Growable(self.rdt, alloc, self)
        self._nodes_host = _Growable(self.ndt, (lambda nbytes, who=None: alloc(nbytes, who, 1)) if hasattr(m, "host_bytes") else alloc, self)
        self._order_host = None           # depth-first order of the first frontier solve's records (device sort)
        self._order_base = 0
        self.stats = SolveStats()
        self.time_kernels = False        # bench.py: CUDA events around every level launch
        self.kernel_ms = []              # (ms, algorithmic bytes, boxes) per level launch
        self.retry_ms = 0.0              # time of launches that overflowed their optimistic buffers
        self._reserve("results", max(256, 4 * self.nsys))
        self._reserve("nodes", max(256, 4 * self.nsys))
        # 2-D, plain arithmetic, tensors that fit shared memory: the round-based frontier pipeline serves the session
        # 2-D, 3-D and 4-D systems in plain arithmetic: the round-based frontier pipelines serve the session
        self.use_frontier = bool(engine.pipeline == 2 and D in (2, 3, 4) and not prm.exact and not prm.trim_policy and self.nsys > 0)
        self.sorted_dt = _lib.sorted_dtype(D)
        self._segments = []               # record ranges in production order: ("level", start, count) / ("frontier", start, count, out)
        self.index_base = 0               # added to every record index the device links with (box-level multi-GPU: rank << 27)
        self.frontier_runner = None       # replaces _run_frontier (distributed.ShardedFrontier)

    # -- growable append-only device arrays ------------------------------------------------
    def _reserve(self, which, want):
        cap = getattr(self, "cap_" + which)
        if want <= cap:
            return
        new_cap = max(want, 2 * cap)
        item = (self.rdt if which == "results" else self.ndt).itemsize
        buf = self.mem.empty(new_cap * item)
        old = getattr(self, "d_" + which)
        used = getattr(self, "n_" + which)
        if old is not None and used:
            self.mem.copy(buf, old, min(used, cap) * item)      # (indices are global: ranges the frontier pipeline produced are holes here)
        setattr(self, "d_" + which, buf)
        setattr(self, "cap_" + which, new_cap)

    def _read_counters(self):
        raw = self.mem.download(self.d_counters, _lib.COUNTER_BYTES).view(np.uint64)
        return dict(zip(_lib.COUNTER_FIELDS, (int(v) for v in raw)))

    def _launch(self, threads, ctas, in_smem, ws, warp=False):
        L = _lib.YrLaunch()
        L.threads, L.ctas, L.ws_in_smem, L.ws_doubles, L.warp_per_box = threads, ctas, int(in_smem), ws, int(warp)
        gws = None
        if not in_smem:
            gws = self.mem.empty(ctas * ws * 8)
            L.gws, L.gws_stride = gws.ptr, ws
        return L, gws

    # -- level 0 -----------------------------------------------------------------------------
    def _init_boxes(self, jobs):
        D, m = self.dim, self.mem
        sysidx = np.ascontiguousarray(jobs["system"], dtype=np.int32)
        nj = len(sysidx)
        total = int(self.sys_total[sysidx].sum())
        keep = [m.upload(sysidx), m.upload(np.ascontiguousarray(jobs["restrict"], dtype=np.int32)),
                m.upload(np.ascontiguousarray(jobs["alpha"], dtype=np.float64)),
                m.upload(np.ascontiguousarray(jobs["beta"], dtype=np.float64)),
                m.upload(np.ascontiguousarray(jobs["iv"], dtype=np.float64)),
                m.upload(np.ascontiguousarray(jobs["level"], dtype=np.int32)),
                m.upload(np.ascontiguousarray(jobs["next_mid"], dtype=np.float64)),
                m.upload(np.ascontiguousarray(jobs["parent"], dtype=np.int32))]
        recs = m.empty(nj * self.box_bytes)
        data = m.empty((total + 2 * nj + 2) * 8)
        threads, ctas, in_smem, ws, _ = self.eng.launch_config(
            D, nj, int(self.sys_total[sysidx].max()), int(self.sys_max_tensor[sysidx].max()),
            int(self.sys_max_extent[sysidx].max()), self.prm.exact, allow_warp=False)
        L, gws = self._launch(threads, ctas, in_smem, ws)
        d = _lib.YrInitDesc()
        d.dim = D
        d.coeffs, d.coeff_off, d.shapes, d.errs = self.d_coeffs.ptr, self.d_off.ptr, self.d_shapes.ptr, self.d_errs.ptr
        (d.job_system, d.job_restrict, d.job_alpha, d.job_beta, d.job_iv, d.job_level, d.job_next_mid,
         d.job_parent) = [b.ptr for b in keep]
        d.njobs = nj
        d.recs_out, d.data_out, d.cap_recs_out, d.cap_data_out = recs.ptr, data.ptr, nj, total + 2 * nj
        d.counters, d.prm, d.launch = self.d_counters.ptr, self.prm, L
        m.zero_(self.d_counters)
        _lib.check(self.lib, self.lib.yr_init_boxes(C.byref(d), m.stream()), "yr_init_boxes")
        c = self._read_counters()
        self.stats.launches += 1
        self.stats.restrict_macs += c["restrict_macs"]
        if c["overflow"]:
            raise RuntimeError("yroots_b200: level-0 buffers too small (internal error)")
        del keep, gws
        return recs, data, c

    # -- the level loop ------------------------------------------------------------------------
    def run_jobs(self, jobs):
        """Solve the given level-0 jobs to completion. Returns (first new result index, count)."""
        D, m, eng = self.dim, self.mem, self.eng
        first_result = self.n_results
        run_frontier = self.frontier_runner or self._run_frontier
        if len(jobs["system"]) == 0:                              # a rank that only receives boxes (sharded frontier)
            return run_frontier(None, None, dict(n_children=0, max_child_extent=0), first_result)
        recs, data, c = self._init_boxes(jobs)
        if self.use_frontier and self.lib.yr_frontier_fits_dim(D, int(c["max_child_extent"]), int(c["max_child_tensor"])):
            return run_frontier(recs, data, c, first_result)
        n_in, data_used = c["n_children"], c["data_used"]
        big_box, big_tensor, big_extent = c["max_child_doubles"], c["max_child_tensor"], c["max_child_extent"]
        fan = 1 << D
        growth = [float(fan), float(fan)]                     # children per box / child data per parent data, last level
        while n_in > 0:
            threads, ctas, in_smem, ws, warp = eng.launch_config(D, n_in, big_box, big_tensor, big_extent, self.prm.exact)
            worst_recs, worst_data = n_in * fan, data_used * fan + 2 * n_in * fan
            if worst_recs * self.box_bytes + worst_data * 8 <= eng.worst_case_budget:
                cap_recs, cap_data = worst_recs, worst_data
            else:
                # the frontier's growth factor changes slowly from level to level: last factor + 30 %
                cap_recs = min(worst_recs, max(4096, int(1.3 * growth[0] * n_in) + 1024))
                cap_data = min(worst_data, max(1 << 16, int(1.3 * growth[1] * data_used) + (1 << 16)))
            pipeline = 1
            work = m.empty(int(self.lib.yr_level_work_bytes(D, n_in)))
            first_c = None
            while True:
                self._reserve("results", self.n_results + n_in)            # a box emits at most one record
                self._reserve("nodes", self.n_nodes + n_in)
                out_recs = m.empty(max(1, cap_recs) * self.box_bytes)
                out_data = m.empty(max(1, cap_data) * 8)
                L, gws = self._launch(threads, ctas, in_smem, ws, warp)
                d = _lib.YrLevelDesc()
                d.dim = D
                d.recs_in, d.data_in, d.n_in = recs.ptr, data.ptr, n_in
                d.recs_out, d.data_out, d.cap_recs_out, d.cap_data_out = out_recs.ptr, out_data.ptr, cap_recs, cap_data
                d.results = self.d_results.ptr + self.n_results * self.rdt.itemsize
                d.cap_results = self.cap_results - self.n_results
                d.nodes = self.d_nodes.ptr + self.n_nodes * self.ndt.itemsize
                d.cap_nodes = self.cap_nodes - self.n_nodes
                d.result_base, d.node_base = self.index_base + self.n_results, self.index_base + self.n_nodes
                d.counters, d.prm, d.launch = self.d_counters.ptr, self.prm, L
                d.pipeline = pipeline
                d.max_box_doubles, d.max_tensor_doubles, d.max_extent = int(big_box), int(big_tensor), int(big_extent)
                if work is not None:
                    d.work, d.work_bytes = work.ptr, work.nbytes
                    d.resume_subdivide = int(first_c is not None)       # after an overflow only the subdivision is redone
                m.zero_(self.d_counters)
                ev = m.timer() if (self.time_kernels and hasattr(m, "timer")) else None
                if ev:
                    ev[0].record()
                _lib.check(self.lib, self.lib.yr_process_level(C.byref(d), m.stream()), "yr_process_level")
                if ev:
                    ev[1].record()
                c = self._read_counters()
                if ev and not c["overflow"]:
                    self.kernel_ms.append((ev[0].elapsed_time(ev[1]), 16 * (first_c or c)["entry_coeffs"], (first_c or c)["boxes"]))
                elif ev:
                    self.retry_ms += ev[0].elapsed_time(ev[1])
                self.stats.launches += 1
                del gws
                if pipeline == 1 and first_c is not None:
                    # the retry only re-ran the subdivision: box statistics and results come from the first pass
                    for k in ("boxes", "zooms", "discard_const", "discard_quad", "discard_zoom", "entry_coeffs", "max_level", "n_results"):
                        c[k] = first_c[k]
                    c["restrict_macs"] += first_c["restrict_macs"]
                if c["no_axis"]:
                    raise _lib.NoAxisLeft("yroots_b200: " + _lib.NO_AXIS_TEXT)
                if not c["overflow"]:
                    break
                self.stats.retries += 1                                      # optimistic buffers were too small
                if pipeline == 1 and first_c is None:
                    first_c = dict(c)
                    first_c["restrict_macs"] = 0                             # (counted again by the re-run's subdivision only partly; keep it simple)
                # grow geometrically towards the worst case (a jump to fan x the level's data can exceed the device:
                # 512 3-D degree-10 systems hold ~30 GB per level)
                cap_recs, cap_data = min(worst_recs, 2 * cap_recs + 4096), min(worst_data, 2 * cap_data + (1 << 16))
                del out_recs, out_data
            del work
            self.stats.add(c)
            self.stats.levels += 1
            self.stats.per_level.append((n_in, c["n_children"], c["n_results"], bool(in_smem), ctas, threads, ws, bool(warp)))
            if c["n_results"] or c["n_nodes"]:
                self._segments.append(("level", self.n_results, int(c["n_results"]), self.n_nodes, int(c["n_nodes"])))
            self.n_results += c["n_results"]
            self.n_nodes += c["n_nodes"]
            growth = [max(0.25, c["n_children"] / n_in), max(0.25, c["data_used"] / max(data_used, 1))]
            recs, data = out_recs, out_data
            n_in, data_used = c["n_children"], c["data_used"]
            big_box, big_tensor, big_extent = c["max_child_doubles"], c["max_child_tensor"], c["max_child_extent"]
            if self.use_frontier and n_in > 0 and self.lib.yr_frontier_fits_dim(D, int(big_extent), int(big_tensor)):
                # large 2-D systems: the first levels run here, the rest of the search on the frontier pipeline
                run_frontier(recs, data, c, first_result)
                break
        return first_result, self.n_results - first_result

    def frontier_desc(self, recs, data, c, max_extent=None):
        d = _lib.YrFrontierDesc()
        d.dim, d.time_kernels = self.dim, int(bool(self.time_kernels))
        if recs is not None:
            d.recs_in, d.data_in = recs.ptr, data.ptr
        d.n_in = int(c["n_children"])
        d.max_extent = int(c["max_child_extent"] if max_extent is None else max_extent)
        d.result_base, d.node_base = self.index_base + self.n_results, self.index_base + self.n_nodes
        d.prm = self.prm
        return d

    def _run_frontier(self, recs, data, c, first_result):
        """2-D systems: the whole search in one call (include/yroots_b200.h: yr_frontier_solve)."""
        m, lib = self.mem, self.lib
        d = self.frontier_desc(recs, data, c)
        out = _lib.YrFrontierOut()
        stream = m.stream()
        _lib.check(lib, lib.yr_frontier_solve(C.byref(d), C.byref(out), stream), "yr_frontier_solve")
        return self.adopt_frontier_out(out, first_result)

    def adopt_frontier_out(self, out, first_result):
        nres, nnod = int(out.n_results), int(out.n_nodes)
        self._pending.append(out)                                 # records stay on the device until results() / nodes()
        self._segments.append(("frontier", self.n_results, nres, self.n_nodes, nnod, out))
        self.n_results += nres
        self.n_nodes += nnod
        cc = {k: int(getattr(out.counters, k)) for k in _lib.COUNTER_FIELDS}
        self.stats.add(cc)
        self.stats.levels += int(out.rounds)
        self.stats.launches += int(out.launches)
        self.stats.rounds = getattr(self.stats, "rounds", 0) + int(out.rounds)
        if self.time_kernels:
            self.kernel_ms.append((float(out.tensor_ms), 16 * cc["entry_coeffs"], cc["boxes"]))
            self.scalar_ms = getattr(self, "scalar_ms", 0.0) + float(out.scalar_ms)
        return first_result, nres

    def unit_jobs(self, systems=None):
        """One job per system: the whole unit cube (ChebyshevSubdivisionSolver.py:1423)."""
        D = self.dim
        idx = np.arange(self.nsys, dtype=np.int32) if systems is None else np.asarray(systems, dtype=np.int32)
        n = len(idx)
        iv = np.empty((n, D, 2))
        iv[:, :, 0], iv[:, :, 1] = -1.0, 1.0
        return dict(system=idx, restrict=np.zeros(n, np.int32), alpha=np.ones((n, D)), beta=np.zeros((n, D)), iv=iv,
                    level=np.zeros(n, np.int32), next_mid=np.full((n, D), FIRST_SPLIT), parent=np.full(n, -1, np.int32))

    # -- results -----------------------------------------------------------------------------
    def _detach_host(self, slot=0):
        if slot == 2:
            if self._order_host is not None:
                self._order_host = self._order_host.copy()
            return
        (self._nodes_host if slot else self._results_host).detach()

    def _fetch_pending(self):
        """Bring every record produced so far to the host, in production (= index) order.  Level-pipeline records
        come from the session's own device arrays, frontier records from the library-owned buffers."""
        lib, stream = self.lib, self.mem.stream()
        while self._segments:
            seg = self._segments.pop(0)
            if seg[0] == "level":
                _, r0, rn, n0, nn = seg
                if rn:
                    item = self.rdt.itemsize
                    self._results_host.tail(rn)[:] = self.mem.download(self.d_results, rn * item, r0 * item)
                if nn:
                    item = self.ndt.itemsize
                    self._nodes_host.tail(nn)[:] = self.mem.download(self.d_nodes, nn * item, n0 * item)
                continue
            _, r0, rn, n0, nn, out = seg
            try:
                for ptr, n, grow in ((out.results, rn, self._results_host), (out.nodes, nn, self._nodes_host)):
                    if n:
                        dst = grow.tail(n)
                        _lib.check(lib, lib.yr_fetch(dst.ctypes.data, ptr, dst.nbytes, stream), "yr_fetch")
                        self.mem.d2h_bytes += dst.nbytes
                if self._order_host is None and rn and out.order:
                    order = np.empty(rn, dtype=self.sorted_dt)
                    _lib.check(lib, lib.yr_fetch(order.ctypes.data, out.order, order.nbytes, stream), "yr_fetch")
                    self.mem.d2h_bytes += order.nbytes
                    self._order_host, self._order_base = order, r0
            finally:
                self._pending.remove(out)
                lib.yr_frontier_release(C.byref(out), stream)

    def release_device(self):
        """Give back everything the session holds on the device: unfetched frontier outputs, the coefficient upload, the
        record arrays.  Host copies and statistics stay.  (A chunked solve calls it after each chunk: the chunks exist
        because the batch does not fit the device at once.)"""
        stream = self.mem.stream()
        for out in self._pending:
            self.lib.yr_frontier_release(C.byref(out), stream)
        self._pending = []
        self._segments = []
        self.d_coeffs = self.d_off = self.d_shapes = self.d_errs = self.d_results = self.d_nodes = None
        self.cap_results = self.cap_nodes = 0

    def __del__(self):
        try:
            for out in self._pending:
                self.lib.yr_frontier_release(C.byref(out), self.mem.stream())
        except Exception:
            pass

    def results(self):
        """All result records so far as a numpy structured array (downloads only what is new)."""
        if self._segments:
            self._fetch_pending()
        return self._results_host.view()

    def nodes(self):
        if self._segments:
            self._fetch_pending()
        return self._nodes_host.view()

    def slim_order(self):
        """The device-sorted (index, system, flags, collector, root) entries of the session's records WITHOUT bringing
        the full result / node records to the host -- possible when one frontier solve produced every record so far
        (the common case: a batch of 2-D systems).  Returns None otherwise.  The full records stay on the device and
        results() / nodes() still fetch them on demand."""
        if self._order_host is not None and self._order_base == 0 and len(self._order_host) == self.n_results:
            return self._order_host
        if len(self._segments) != 1 or self._segments[0][0] != "frontier":
            return None
        _, r0, rn, n0, nn, out = self._segments[0]
        if r0 != 0 or rn != self.n_results or not rn or not out.order:
            return None
        alloc = getattr(self.mem, "host_bytes", None)
        nbytes = rn * self.sorted_dt.itemsize
        raw = alloc(nbytes, self, 2) if alloc is not None else np.empty(nbytes, dtype=np.uint8)
        _lib.check(self.lib, self.lib.yr_fetch(raw.ctypes.data, out.order, nbytes, self.mem.stream()), "yr_fetch")
        self.mem.d2h_bytes += nbytes
        self._order_raw = raw
        self._order_host, self._order_base = raw.view(self.sorted_dt), 0
        return self._order_host

    def sparse_records(self, systems, n_results_hint=None):
        """Bring the full result and node records of `systems` ONLY to the host (one device gather each:
        yr_gather_by_system), as tables of the usual size and indexing whose other records read as zeros.  For a
        session that slim_order() serves: a batch where a few systems need the tree bookkeeping of the host post-pass
        (collectors, merges) no longer pays the device -> host copy of every record.  `n_results_hint`: how many result
        records those systems have, when the caller knows.  Returns False, and fetches nothing, when the session is
        not in that state; results() / nodes() then bring everything."""
        if len(self._segments) != 1 or self._segments[0][0] != "frontier":
            return False
        _, r0, rn, n0, nn, out = self._segments[0]
        if r0 != 0 or n0 != 0 or rn != self.n_results:
            return False
        m, lib = self.mem, self.lib
        stream = m.stream()
        mark = np.zeros(self.nsys, dtype=np.uint8)
        mark[np.asarray(sorted(systems), dtype=np.int64)] = 1
        d_mark = m.upload(mark)
        d_n = m.zeros(16)
        tables = ((out.results, rn, self.rdt, self._results_host, n_results_hint), (out.nodes, nn, self.ndt, self._nodes_host, None))
        caps = [int(hint) if hint is not None else 4 * int(n_results_hint or 0) + 4096 for *_, hint in tables]
        while True:
            # one gather per table into [records | indices]; the counters tell whether the guessed capacity was enough
            m.zero_(d_n)
            bufs = []
            for t, (ptr, n, dt, grow, _) in enumerate(tables):
                item, cap = dt.itemsize, min(caps[t], n)
                caps[t] = cap
                buf = m.empty(cap * (item + 4))
                bufs.append(buf)
                _lib.check(lib, lib.yr_gather_by_system(ptr, n, item, dt.fields["system"][1], d_mark.ptr, self.nsys,
                                                        buf.ptr + cap * item, buf.ptr, cap, d_n.ptr + 8 * t, stream),
                           "yr_gather_by_system")
            counts = [int(v) for v in m.download(d_n, 16).view(np.uint64)]
            if all(c <= cap for c, cap in zip(counts, caps)):
                break
            caps = [max(c, cap) for c, cap in zip(counts, caps)]
        for t, (ptr, n, dt, grow, _) in enumerate(tables):
            item, cap, cnt = dt.itemsize, caps[t], counts[t]
            table, lease = self.eng.sparse_pool.borrow(t, n + max(4096, n // 8), dt, self)      # (room for re-solved boxes)
            if cnt:
                raw = m.download(bufs[t], cap * (item + 4))
                idx = raw[cap * item:cap * item + 4 * cnt].view(np.uint32).copy()
                table.view(np.uint8).reshape(len(table), item)[idx] = raw[:cnt * item].reshape(cnt, item)
                lease["idx"].append(idx)
            grow.adopt(table, n, lease)
        self._segments.clear()
        self._pending.remove(out)
        lib.yr_frontier_release(C.byref(out), stream)
        self.sparse_systems = set(int(v) for v in systems)
        return True

    def dfs_order(self):
        """(order, base, n): base + order["index"] puts records [base, base + n) in (system, depth-first path) order,
        sorted on the device by yr_frontier_solve (order also carries system and root of every record, gathered on
        the device); (None, 0, 0) when the session has none."""
        self.results()
        if self._order_host is None:
            return None, 0, 0
        return self._order_host, self._order_base, len(self._order_host)
