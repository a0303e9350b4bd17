"""Multi-GPU: the batch shards by system, one process per GPU, no data-path collective.

Systems of a batch never interact, so rank r solves systems r, r+W, r+2W, ... on its own
device frontier.  The only exchange is at the end: an all-gather of the per-system root counts
followed by an all-gather of the (padded) roots, so that every rank returns the full answer --
NCCL over NVLink on GPUs, gloo in the CPU tests.  No floating-point reduction is involved, so the
result does not depend on the number of ranks.
"""
import ctypes as C

import numpy as np


def shard(n_items, rank, world):
    """Indices of the systems rank `rank` owns (round-robin)."""
    return np.arange(rank, n_items, world)


_pinned_pool = None         # page-locked result buffers of all_gather_roots (engine.HostPool with a pinned allocator)
from .engine import HostPool as _HostPool          # noqa: E402  (pooled result buffers)


def all_gather_roots(local_roots, local_ids, n_systems, dim, group=None, device=None):
    """Every rank contributes {system id -> roots}; returns the list of roots for all systems."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    counts = torch.zeros(n_systems, dtype=torch.int64, device=dev)
    if len(local_ids):
        counts[torch.as_tensor(np.asarray(local_ids), device=dev)] = torch.as_tensor(
            np.array([len(r) for r in local_roots], dtype=np.int64), device=dev)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)            # integer counts, owners are disjoint
    counts_h = counts.cpu().numpy()
    per_rank = np.array([int(counts_h[shard(n_systems, r, world)].sum()) for r in range(world)])
    cap = int(per_rank.max()) if world else 0
    mine = torch.zeros((max(cap, 1), dim), dtype=torch.float64, device=dev)
    if len(local_roots) and sum(len(r) for r in local_roots):
        flat = np.concatenate([np.asarray(r, dtype=np.float64).reshape(-1, dim) for r in local_roots])
        mine[:len(flat)] = torch.as_tensor(flat, device=dev)
    # one receive buffer, one device -> host copy (page-locked on GPUs), per-system views instead of copies: the cost
    # of handing every rank the whole answer stays a small part of the solve as the number of ranks grows
    big = torch.empty((world, max(cap, 1), dim), dtype=torch.float64, device=dev)
    dist.all_gather([big[r] for r in range(world)], mine, group=group)
    if big.is_cuda:
        # device -> a page-locked result buffer from the pool, directly: no staging copy, and no pinned allocation per call
        # (a buffer returns to the pool when the caller has dropped every root array of an earlier call)
        global _pinned_pool
        if _pinned_pool is None:
            def _pinned(nbytes):
                return torch.empty(int(nbytes), dtype=torch.uint8, pin_memory=True).numpy()   # (the array keeps the tensor alive)
            _pinned_pool = _HostPool(alloc=_pinned)
        nbytes = big.numel() * 8
        blocks = _pinned_pool.take(nbytes).view(np.float64).reshape(tuple(big.shape))
        torch.from_numpy(blocks).copy_(big, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    else:
        blocks = big.numpy()
    out = [None] * n_systems
    for r in range(world):
        ids = shard(n_systems, r, world)
        if not len(ids):
            continue
        ends = np.cumsum(counts_h[ids])
        starts = ends - counts_h[ids]
        blk = blocks[r]
        for s_id, lo, hi in zip(ids.tolist(), starts.tolist(), ends.tolist()):     # (plain slices: np.split costs 3 us a piece)
            out[s_id] = blk[lo:hi]
    return out


def solve_batch_sharded(systems, errors, group=None, engine=None, **kw):
    """`solve_batch` over all ranks of the process group; every rank gets all roots."""
    import torch.distributed as dist
    from .ChebyshevSubdivisionSolver import solve_batch
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ids = shard(len(systems), rank, world)
    local = solve_batch([systems[i] for i in ids], [errors[i] for i in ids], engine=engine, **kw) if len(ids) else []
    return all_gather_roots(local, ids, len(systems), len(systems[0]), group=group)


# =====================================================================================================================
# Box-level mode: ONE frontier sharded over the ranks, with periodic load rebalancing.
#
# Sibling boxes of a subdivision are independent (/root/reference/yroots/ChebyshevSubdivisionSolver.py:1314-1317), so
# the queued boxes of a frontier may run on any GPU.  Every rank drives its own device frontier in rounds
# (yr_frontier_begin / _advance); every `every` rounds the ranks all-gather their queue lengths, derive the same
# transfer plan, and donors ship surplus boxes -- BoxRec + Work + both tensors each, packed by one kernel straight into
# the send buffer (yr_frontier_export) -- to receivers (send/recv: NCCL over NVLink on GPUs, gloo in the CPU tests),
# which adopt them (yr_frontier_import).  No floating-point value is combined across ranks and every box carries its
# depth-first path, so roots, their order and the box count do not depend on the number of ranks.  At the end the
# (few) result and node records travel to rank 0, which runs the tree bookkeeping (postpass) and broadcasts the roots.
# =====================================================================================================================
RANK_SHIFT = 27                    # rank r links its records with indices (r << 27) + local index


def rebalance_plan(counts, min_move=256, tolerance=1.25):
    """Deterministic transfer list [(src, dst, n)] that evens out `counts` (queued boxes per rank); empty when the
    load is already within `tolerance` of the mean or too small to be worth moving."""
    counts = [int(c) for c in counts]
    world, total = len(counts), sum(counts)
    if world < 2 or total < world * min_move:
        return []
    mean = total / world
    if max(counts) <= tolerance * mean:
        return []
    target = [total // world + (1 if r < total % world else 0) for r in range(world)]
    give = [[r, counts[r] - target[r]] for r in range(world) if counts[r] > target[r]]
    take = [[r, target[r] - counts[r]] for r in range(world) if counts[r] < target[r]]
    plan, gi, ti = [], 0, 0
    while gi < len(give) and ti < len(take):
        n = min(give[gi][1], take[ti][1])
        if n >= min_move:
            plan.append((give[gi][0], take[ti][0], n))
        give[gi][1] -= n
        take[ti][1] -= n
        if give[gi][1] == 0:
            gi += 1
        if take[ti][1] == 0:
            ti += 1
    return plan


def _bytes_tensor(mem, nbytes):
    """(torch uint8 tensor, device pointer) of nbytes in the engine's memory space."""
    import torch
    buf = mem.empty(nbytes)
    owner = buf.owner
    t = owner.view(torch.uint8) if isinstance(owner, torch.Tensor) else torch.from_numpy(owner.view(np.uint8))
    return t, buf.ptr, buf


class ShardedFrontier:
    """Runs a SolveSession's 2-D frontier phase together with the other ranks of `group` (see above)."""

    def __init__(self, sess, group=None, every=2, min_move=256, tolerance=1.25):
        import torch.distributed as dist
        self.sess, self.group = sess, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.every, self.min_move, self.tolerance = every, min_move, tolerance
        self.rebalances = 0
        self.boxes_moved = 0
        self.bytes_moved = 0
        self.log = []                       # (round, counts, plan)
        sess.index_base = self.rank << RANK_SHIFT
        sess.frontier_runner = self.run

    def _device(self):
        import torch
        t = getattr(self.sess.mem, "device", None)
        return t if t is not None else torch.device("cpu")

    def _all_gather_ints(self, values):
        import torch
        import torch.distributed as dist
        mine = torch.tensor(values, dtype=torch.int64, device=self._device())
        out = torch.empty((self.world, len(values)), dtype=torch.int64, device=self._device())
        dist.all_gather_into_tensor(out, mine, group=self.group) if out.is_cuda else dist.all_gather(list(out.unbind(0)), mine, group=self.group)
        return out.cpu().numpy()

    def run(self, recs, data, c, first_result):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib
        sess, lib, mem = self.sess, self.sess.lib, self.sess.mem
        stream = mem.stream()
        # every rank must build its subdivision tables for the largest extent any rank holds
        ext = int(self._all_gather_ints([int(c["max_child_extent"])]).max())
        d = sess.frontier_desc(recs, data, c, max_extent=max(ext, 4))
        _lib.check(lib, lib.yr_frontier_begin(C.byref(d), stream), "yr_frontier_begin")
        prog = _lib.YrFrontierProgress()
        dev = self._device()
        while True:
            _lib.check(lib, lib.yr_frontier_advance(self.every, C.byref(prog), stream), "yr_frontier_advance")
            counts = self._all_gather_ints([int(prog.pending)])[:, 0]
            if int(counts.sum()) == 0:
                break
            plan = rebalance_plan(counts, self.min_move, self.tolerance)
            self.log.append((int(prog.rounds), counts.tolist(), plan))
            if plan:
                self.rebalances += 1
            for src, dst, n in plan:                       # same order on every rank: sends and receives pair up
                if self.rank == src:
                    cap = int(lib.yr_frontier_export_bound(n))
                    t, ptr, keep = _bytes_tensor(mem, cap)
                    taken, nbytes = C.c_uint64(0), C.c_uint64(0)
                    _lib.check(lib, lib.yr_frontier_export(n, ptr, cap, C.byref(taken), C.byref(nbytes), stream), "yr_frontier_export")
                    mem.synchronize()
                    dist.send(torch.tensor([taken.value, nbytes.value], dtype=torch.int64, device=dev), dst, group=self.group)
                    if nbytes.value:
                        dist.send(t[:nbytes.value], dst, group=self.group)
                    self.boxes_moved += int(taken.value)
                    self.bytes_moved += int(nbytes.value)
                    del t, keep
                elif self.rank == dst:
                    head = torch.empty(2, dtype=torch.int64, device=dev)
                    dist.recv(head, src, group=self.group)
                    taken, nbytes = (int(v) for v in head.cpu().tolist())
                    if nbytes:
                        t, ptr, keep = _bytes_tensor(mem, nbytes)
                        dist.recv(t[:nbytes], src, group=self.group)
                        if t.is_cuda:
                            torch.cuda.current_stream().synchronize()
                        _lib.check(lib, lib.yr_frontier_import(ptr, nbytes, stream), "yr_frontier_import")
                        mem.synchronize()
                        del t, keep
        out = _lib.YrFrontierOut()
        _lib.check(lib, lib.yr_frontier_finish(C.byref(out), stream), "yr_frontier_finish")
        return sess.adopt_frontier_out(out, first_result)


def _remap(ids, res_base, counts):
    """Global record ids (rank << RANK_SHIFT) + local -> positions in rank 0's concatenated table."""
    ids = np.asarray(ids, dtype=np.int64)
    out = ids.copy()
    ok = ids >= 0
    r = ids[ok] >> RANK_SHIFT
    local = ids[ok] & ((1 << RANK_SHIFT) - 1)
    out[ok] = np.asarray(res_base, dtype=np.int64)[r] + local
    return out.astype(np.int32)


def solve_sharded(systems, errors, group=None, engine=None, every=2, min_move=256, tolerance=1.25, returnBoundingBoxes=False,
                  return_stats=False, start="round_robin", **kw):
    """solve_batch with the FRONTIER (not the list of systems) spread over the ranks of `group`: for one many-root
    system or a few large ones.  Rank r starts the systems r, r+W, ... (start="rank0": rank 0 starts everything and the
    rebalancer spreads the frontier); queued boxes then move between ranks whenever the queue lengths drift apart.  Every
    rank returns all roots (2-D systems, plain arithmetic)."""
    import torch
    import torch.distributed as dist
    from . import postpass
    from .runtime import get_engine
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world > (1 << (31 - RANK_SHIFT)):
        raise ValueError("solve_sharded supports up to 16 ranks")
    eng = get_engine() if engine is None else engine
    if len(systems) == 0:
        return ([], []) if returnBoundingBoxes else []
    if len(systems[0]) != 2:
        raise ValueError("solve_sharded: the sharded frontier serves 2-D systems (others shard by system: solve_batch_sharded)")
    sess = eng.session(systems, errors, use_fma=int(bool(kw.pop("use_fma", False))),
                       constant_check=int(bool(kw.pop("constant_check", True))),
                       low_dim_quad_check=int(bool(kw.pop("low_dim_quadratic_check", True))),
                       all_dim_quad_check=int(bool(kw.pop("all_dim_quadratic_check", False))))
    if kw:
        raise TypeError(f"unknown arguments {sorted(kw)}")
    import time
    sf = ShardedFrontier(sess, group, every, min_move, tolerance)
    t_start = time.perf_counter()
    mine = shard(len(systems), rank, world) if start == "round_robin" else (np.arange(len(systems)) if rank == 0 else np.arange(0))
    sess.run_jobs(sess.unit_jobs(mine))
    sess.mem.synchronize()
    t_solved = time.perf_counter()
    # ---- records of every rank to rank 0 ----
    from . import _lib
    lib, mem = sess.lib, sess.mem
    dev = sf._device()
    rdt, ndt = sess.rdt, sess.ndt
    seg = sess._segments[0] if len(sess._segments) == 1 and sess._segments[0][0] == "frontier" and sess._segments[0][1] == 0 else None
    on_device = bool(sf._all_gather_ints([int(seg is not None and hasattr(lib, "yr_frontier_adopt"))]).min())
    if on_device:
        # Every rank's records are still on its device (one frontier solve each): they travel device to device, rank 0
        # concatenates them, has the links remapped and the reporting order computed there (yr_frontier_adopt) and goes on
        # exactly like a single-GPU solve: device-sorted entries, full records only for the systems that need them.
        _, _, rn, _, nn, out = seg
        sizes = sf._all_gather_ints([rn, nn])
        stream = mem.stream()
        if rank == 0:
            tot_r, tot_n = int(sizes[:, 0].sum()), int(sizes[:, 1].sum())
            res_base = np.r_[0, np.cumsum(sizes[:, 0])[:-1]].astype(np.uint64)
            node_base = np.r_[0, np.cumsum(sizes[:, 1])[:-1]].astype(np.uint64)
            bufs = []
            for col, (ptr, dt, tot, base) in enumerate(((out.results, rdt, tot_r, res_base), (out.nodes, ndt, tot_n, node_base))):
                t, tptr, keep = _bytes_tensor(mem, max(1, tot * dt.itemsize))
                bufs.append((t, tptr, keep))
                if int(sizes[0, col]):
                    _lib.check(lib, lib.yr_device_copy(tptr, ptr, int(sizes[0, col]) * dt.itemsize, stream), "yr_device_copy")
                mem.synchronize()
                for r in range(1, world):
                    n = int(sizes[r, col])
                    if n:
                        lo = int(base[r]) * dt.itemsize
                        dist.recv(t[lo:lo + n * dt.itemsize], r, group=group)
            if dev.type == "cuda":
                torch.cuda.current_stream().synchronize()
            merged = _lib.YrFrontierOut()
            _lib.check(lib, lib.yr_frontier_adopt(2, bufs[0][1], tot_r, bufs[1][1], tot_n,
                                                  res_base.ctypes.data_as(C.POINTER(C.c_uint64)), node_base.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                  world, RANK_SHIFT, C.byref(merged), stream), "yr_frontier_adopt")
            mem.synchronize()
            del bufs
            sess._pending.remove(out)
            lib.yr_frontier_release(C.byref(out), stream)
            sess._pending.append(merged)
            sess._segments = [("frontier", 0, tot_r, 0, tot_n, merged)]
            sess.n_results, sess.n_nodes = tot_r, tot_n
            sess._order_host = None
        else:
            for ptr, n, dt in ((out.results, rn, rdt), (out.nodes, nn, ndt)):
                if n:
                    t, tptr, keep = _bytes_tensor(mem, n * dt.itemsize)
                    _lib.check(lib, lib.yr_device_copy(tptr, ptr, n * dt.itemsize, stream), "yr_device_copy")
                    mem.synchronize()
                    dist.send(t[:n * dt.itemsize], 0, group=group)
                    del t, keep
            sess._pending.remove(out)
            lib.yr_frontier_release(C.byref(out), stream)
            sess._segments = []
    else:
        res, nodes = sess.results(), sess.nodes()
        sizes = sf._all_gather_ints([len(res), len(nodes)])
        if rank == 0:
            res_base = np.r_[0, np.cumsum(sizes[:, 0])[:-1]]
            node_base = np.r_[0, np.cumsum(sizes[:, 1])[:-1]]
            for r in range(1, world):
                for which, n, dt in (("res", int(sizes[r, 0]), rdt), ("nodes", int(sizes[r, 1]), ndt)):
                    if not n:
                        continue
                    t = torch.empty(n * dt.itemsize, dtype=torch.uint8, device=dev)
                    dist.recv(t, r, group=group)
                    grow = sess._results_host if which == "res" else sess._nodes_host
                    grow.tail(n)[:] = t.cpu().numpy()
            sess.n_results, sess.n_nodes = int(sizes[:, 0].sum()), int(sizes[:, 1].sum())
            sess._order_host = None                            # the device order covers rank 0's records only
            res, nodes = sess._results_host.view(), sess._nodes_host.view()
            res["parent_node"] = _remap(res["parent_node"], node_base, sizes[:, 1])
            res["collector"] = _remap(res["collector"], res_base, sizes[:, 0])
            nodes["parent_node"] = _remap(nodes["parent_node"], node_base, sizes[:, 1])
        else:
            for arr in (res, nodes):
                if len(arr):
                    t = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy()).to(dev)
                    dist.send(t, 0, group=group)
    if rank == 0:
        sess.index_base, sess.frontier_runner = 0, None    # merge re-solves run on this rank alone
        res, keep, dup, extra = postpass.assemble(sess, need_records=returnBoundingBoxes)
        roots, boxes = postpass.per_system_roots(res, keep, dup, extra, sess.nsys, sess.dim, want_boxes=returnBoundingBoxes,
                                                 sorted_fields=getattr(sess, "sorted_fields", None),
                                                 late_by_system=getattr(sess, "late_by_system", None),
                                                 redo_systems=getattr(sess, "redo_systems", None), alive=getattr(sess, "alive", None))
        payload = [roots, [[(b.finalInterval, b.possibleDuplicateRoots, b.possibleExtraRoot) for b in bs] for bs in boxes] if returnBoundingBoxes else None]
    else:
        payload = [None, None]
    t_assembled = time.perf_counter()
    stats = sf._all_gather_ints([int(sess.stats.boxes), sf.boxes_moved, sf.bytes_moved, sf.rebalances,
                                 int(1e6 * (t_solved - t_start)), int(1e6 * (t_assembled - t_solved))])
    dist.broadcast_object_list(payload, src=0, group=group, device=dev if dev.type == "cuda" else None)
    roots, boxes = payload
    out = (roots, boxes) if returnBoundingBoxes else roots
    if return_stats:
        info = dict(boxes=int(stats[:, 0].sum()), boxes_per_rank=stats[:, 0].tolist(), boxes_moved=int(stats[:, 1].sum()),
                    bytes_moved=int(stats[:, 2].sum()), rebalances=int(stats[:, 3].max()), log=sf.log,
                    solve_ms=float(stats[:, 4].max()) / 1e3,          # level-0 boxes -> last frontier round, slowest rank
                    assemble_ms=float(stats[:, 5].max()) / 1e3)       # records to rank 0, tree bookkeeping there
        return out, info
    return out
