"""Tree-structured bookkeeping the reference does while its recursion unwinds.

The device frontier is breadth first and emits flat result records; two pieces of
solvePolyRecursive are defined on the recursion tree and act only on the (small) set of
result boxes, so they run here on the host:

 A. final-step subdivision (ChebyshevSubdivisionSolver.py:1266-1298): a box that has to be
    subdivided *inside* its final step reports itself; the boxes found below it are
    de-duplicated by the exact lower corner of their interval and become its
    `possibleDuplicateRoots`; if nothing is found it is kept with `possibleExtraRoot`.
 B. merging of exterior boxes (:1318-1385): result boxes of different children that touch or
    overlap are replaced by the solve of their bounding box, started from the lowest common
    ancestor's level.  The re-solve restricts the *top-level* tensors of the system to the
    merged interval on the device (yr_init_boxes with job_restrict=1).

Both are rare (2 of the 27 Chebfun2 systems); everything is vectorised so that
batches with millions of ordinary results pay only a few numpy passes.
"""
import warnings

import numpy as np

RES_FINAL, RES_EXTRA, RES_COLLECTOR, RES_TOUCH, RES_POINT = 1, 2, 4, 8, 16


class ResultBox:
    """What the reference returns per root box (a TrackedInterval, :376-576), read-only."""

    def __init__(self, rec, dup_roots=None, extra=False):
        self.finalInterval = np.array(rec["box"])
        self.interval = np.array(rec["comb_iv"])
        self.topInterval = np.array([[-1., 1.]] * len(self.interval))
        self.root = np.array(rec["root"])
        self.possibleDuplicateRoots = [] if dup_roots is None else dup_roots
        self.possibleExtraRoot = bool(extra)
        self.finalStep = bool(rec["flags"] & RES_FINAL)
        self.ndim = len(self.interval)
        self.system = int(rec["system"])
        self.level = int(rec["level"])

    def getFinalInterval(self):
        return self.finalInterval

    def getFinalPoint(self):
        return self.root

    def finalDimSize(self):
        return self.finalInterval[:, 1] - self.finalInterval[:, 0]

    def dimSize(self):
        return self.interval[:, 1] - self.interval[:, 0]

    def getIntervalForCombining(self):
        return self.interval

    def __contains__(self, point):
        return bool(np.all(point >= self.interval[:, 0]) and np.all(point <= self.interval[:, 1]))

    def __repr__(self):
        return str(self.finalInterval)


def _path_order(res, idx):
    """Records by (system, depth-first path).  Paths start at the level-0 job, so records of several re-solved merged
    boxes of one system tie; the root coordinates break the tie (the record index would depend on the order in which
    the device happened to emit them)."""
    root = res["root"][idx]
    keys = tuple(root[:, d] for d in range(root.shape[1] - 1, -1, -1))
    return idx[np.lexsort(keys + (res["path_lo"][idx], res["path_hi"][idx], res["system"][idx]))]


def resolve_collectors(res, start=0, stop=None, known=None):
    """Step A on records [start, stop).  Record links (`collector`) are global indices into `res`.  `known` =
    (collector records, member records) when the caller already has them (from the device-gathered entries).

    Returns (alive mask for that range, {global record index: list of roots}, set of extra-root records).
    """
    stop = len(res) if stop is None else stop
    alive = np.ones(stop - start, dtype=bool)
    dup = {}
    extra = set()
    if known is not None:
        coll, member = known
    else:
        rng = np.arange(start, stop)
        coll = rng[(res["flags"][start:stop] & RES_COLLECTOR) != 0]
        member = rng[res["collector"][start:stop] >= 0]
    if len(coll) == 0 and len(member) == 0:
        return alive, dup, extra
    alive[member - start] = False              # everything below a collector is absorbed by it
    kids = {}
    for i in _path_order(res, member):         # depth-first order == the reference's visiting order
        kids.setdefault(int(res["collector"][i]), []).append(int(i))
    for c in coll[np.argsort(-res["level"][coll], kind="stable")]:     # deepest collectors first
        c = int(c)
        seen, roots = set(), []
        for i in kids.get(c, []):
            key = tuple(res["iv_lo"][i])
            if key in seen:
                continue
            seen.add(key)
            if dup.get(i):
                roots += dup[i]
            else:
                roots.append(np.array(res["root"][i]))
        if roots:
            dup[c] = roots
        else:
            extra.add(c)
    return alive, dup, extra


def _overlap(a, b):
    """TrackedInterval.overlapsWith :548-556 (closed intervals: touching counts)."""
    return bool(np.all(a[:, 0] <= b[:, 1]) and np.all(b[:, 0] <= a[:, 1]))


def _ancestors(nodes, start):
    chain = []
    k = int(start)
    while k >= 0:
        chain.append(k)
        k = int(nodes["parent_node"][k])
    return chain


def _exterior_at(res, nodes, i, node, chain, resolved_under):
    """Is live record i in `resultExterior` when the reference's recursion unwinds to subdividing node `node`?
    (`chain` = ancestors of i, innermost first.)  A box travels upwards as long as it shares an endpoint with the
    cell of every call it leaves (:1381, isExteriorInterval); boxes a merge re-solved are passed on unchecked by
    the node that re-solved them (:1376-1380)."""
    if node < 0:
        pos = len(chain)
    else:
        pos = chain.index(node)
    if not (res["flags"][i] & RES_TOUCH):
        return False
    iv = res["comb_iv"][i]
    for k in chain[:pos]:                           # cells left on the way up: the subdividing calls below `node`
        if resolved_under.get(int(i)) == k:         # re-solved by the merge at k: passed on unchecked
            continue
        c = nodes["iv"][k]
        if not (np.any(iv[:, 0] == c[:, 0]) or np.any(iv[:, 1] == c[:, 1])):
            return False
    return True


def merge_exterior(session, res, alive, dup, extra, max_rounds=64, cand=None):
    """Step B.  May run more device jobs through `session`; returns the updated (res, alive, dup, extra).
    `cand`: the live records flagged RES_TOUCH, when the caller already knows them.

    Per system and round: the deepest subdividing node two overlapping exterior boxes have in common is where the
    unwinding recursion meets them first.  There the reference merges EVERY transitively overlapping exterior box
    of that node's subtree into growing bounding boxes (:1323-1367) and re-solves each merged box once (:1369-1380);
    three or more boxes that meet in one point (a root on a corner of the subdivision grid) therefore become one
    re-solve, not several pairs."""
    rounds = 0
    # candidates: live records whose reported interval touches the cell they were created as; the set is small
    # and is maintained incrementally (one pass over the records, not one per merge)
    if cand is None:
        cand = np.nonzero(alive & ((res["flags"] & RES_TOUCH) != 0))[0]
    cand_set = set(np.asarray(cand).tolist())
    resolved_under = getattr(session, "_resolved_under", None)
    if resolved_under is None:
        resolved_under = session._resolved_under = {}
    dim = res["comb_iv"].shape[1]
    while rounds < max_rounds:
        rounds += 1
        if len(cand_set) < 2:
            break
        cand = np.fromiter(sorted(cand_set), dtype=np.int64, count=len(cand_set))
        order = cand[np.argsort(res["system"][cand], kind="stable")]
        sysv = res["system"][order]
        starts = np.nonzero(np.r_[True, sysv[1:] != sysv[:-1], True])[0]
        nodes = None
        jobs, progress = [], False
        for s, e in zip(starts[:-1], starts[1:]):
            grp = order[s:e]
            if len(grp) < 2:
                continue
            iv = res["comb_iv"][grp]
            pairs = [(x, y) for x in range(len(grp)) for y in range(x + 1, len(grp)) if _overlap(iv[x], iv[y])]
            if not pairs:
                continue
            if nodes is None:
                nodes = session.nodes()
            chains = [_ancestors(nodes, res["parent_node"][g]) for g in grp]
            # the node the unwinding recursion merges at first: the deepest common ancestor over all overlapping pairs
            best = None
            for x, y in pairs:
                sy = set(chains[y])
                lca = next((k for k in chains[x] if k in sy), -1)
                lvl = int(nodes["level"][lca]) if lca >= 0 else 0
                if best is None or lvl > best[0]:
                    best = (lvl, lca)
            lvl, lca = best
            if lca >= 0:
                anc_iv, anc_mid, anc_level = nodes["iv"][lca], nodes["next_mid"][lca], int(nodes["level"][lca])
            else:                                   # the boxes come from different level-0 jobs of one system
                anc_iv = np.array([[-1., 1.]] * dim)
                anc_mid, anc_level = np.full(dim, 0.0394555475981047), 0
            # resultExterior of that call, in the reference's (depth-first) order
            members = [x for x in range(len(grp)) if (lca < 0 or lca in chains[x])
                       and _exterior_at(res, nodes, int(grp[x]), lca, chains[x], resolved_under)]
            members = [int(g) for g in _path_order(res, grp[members])]
            # the reference's double loop :1326-1367 on (interval, member records)
            work = [(res["comb_iv"][g].copy(), [g]) for g in members]
            i1 = 0
            while i1 < len(work):
                i2 = i1 + 1
                while i2 < len(work):
                    if _overlap(work[i1][0], work[i2][0]):
                        a, b = work[i1], work[i2]
                        box = np.stack([np.minimum(a[0][:, 0], b[0][:, 0]), np.maximum(a[0][:, 1], b[0][:, 1])], axis=1)
                        del work[i2]
                        del work[i1]
                        work.append((box, a[1] + b[1]))
                        i2 = i1 + 1
                    else:
                        i2 += 1
                i1 += 1
            sysid = int(res["system"][grp[0]])
            for box, recs in work:
                if len(recs) < 2:
                    continue
                lo, hi = box[:, 0], box[:, 1]
                progress = True
                for g in recs:
                    alive[g] = False
                    cand_set.discard(g)
                session.stats_merges = getattr(session, "stats_merges", 0) + 1
                if np.array_equal(lo, anc_iv[:, 0]) and np.array_equal(hi, anc_iv[:, 1]):
                    # :1370-1371 the merged box is the ancestor's whole box: keep it, do not re-solve
                    keep = recs[0]
                    res["comb_iv"][keep][:, 0], res["comb_iv"][keep][:, 1] = lo, hi
                    res["box"][keep][:, 0] = np.min(res["box"][recs][:, :, 0], axis=0)
                    res["box"][keep][:, 1] = np.max(res["box"][recs][:, :, 1], axis=0)
                    res["root"][keep] = (res["box"][keep][:, 0] + res["box"][keep][:, 1]) / 2
                    session.patched_roots = getattr(session, "patched_roots", set()) | {int(keep)}
                    res["flags"][keep] &= ~np.uint32(RES_TOUCH)
                    alive[keep] = True
                    continue
                jobs.append((sysid, lo.copy(), hi.copy(), anc_level, np.array(anc_mid), lca))
        if not jobs:
            if not progress:
                break                                # no overlapping pair anywhere (or none the reference would merge)
            continue
        n_old_nodes = len(nodes)
        los, his = np.array([q[1] for q in jobs]), np.array([q[2] for q in jobs])
        job = dict(system=np.array([q[0] for q in jobs], np.int32), restrict=np.ones(len(jobs), np.int32),
                   alpha=(his - los) / 2, beta=(his + los) / 2, iv=np.stack([los, his], axis=2),
                   level=np.array([q[3] for q in jobs], np.int32), next_mid=np.array([q[4] for q in jobs]),
                   parent=np.array([q[5] for q in jobs], np.int32))
        first, count = session.run_jobs(job)
        res = session.results()
        grow = len(res) - len(alive)
        alive = np.r_[alive, np.ones(grow, dtype=bool)]
        sub_alive, sub_dup, sub_extra = resolve_collectors(res, first, first + count)
        alive[first:first + count] = sub_alive
        dup.update(sub_dup)
        extra |= sub_extra
        new = np.arange(first, first + count)
        fresh = new[sub_alive & ((res["flags"][first:first + count] & RES_TOUCH) != 0)].tolist()
        cand_set.update(fresh)
        if fresh:
            nodes = session.nodes()
            for g in fresh:                          # the node whose merge re-solved this box passes it on unchecked
                k = int(res["parent_node"][g])
                while k >= n_old_nodes:
                    k = int(nodes["parent_node"][k])
                resolved_under[int(g)] = k
    return res, alive, dup, extra


def assemble(session, merge=True, need_records=False):
    """Turn the session's result records into per-system (roots, boxes) in depth-first order.  `need_records`: the
    caller wants the full records (bounding boxes), not only the roots.

    Returns (res, keep, dup, extra) for per_system_roots and leaves on the session
      sorted_fields   (system, root) of `keep` in reporting order when the device gathered them,
      late_by_system  live records outside `keep` (first levels of a large system, re-solved merged boxes),
      redo_systems    systems that lost records of `keep` to a merge or a collector (assembled one by one).
    """
    slim = session.slim_order() if hasattr(session, "slim_order") else None
    session.redo_systems = set()
    if slim is not None:
        # one frontier solve produced every record: the device-gathered entries (reporting order) carry what the
        # common case needs; the full records come to the host only for bounding boxes, collectors or merges
        fl = np.ascontiguousarray(slim["flags"])                  # (the entries are 32-byte records: strided fields)
        sysv = np.ascontiguousarray(slim["system"])
        is_coll, is_member = (fl & RES_COLLECTOR) != 0, slim["collector"] >= 0
        special = bool(is_coll.any() or is_member.any())
        touch_pos = np.nonzero((fl & RES_TOUCH) != 0)[0]
        touch_sys = sysv[touch_pos]
        may_merge = merge and len(touch_pos) >= 2 and len(np.unique(touch_sys)) < len(touch_sys)
        keep = slim["index"]
        bounds = np.searchsorted(sysv, np.arange(session.nsys + 1))
        # the roots handed to the caller: a pooled buffer (a fresh 11 MB array costs its first-touch page faults every call)
        from .engine import result_pool
        roots_all = result_pool.take(slim["root"].size * 8).view(np.float64).reshape(slim["root"].shape)
        roots_all[...] = slim["root"]
        session.sorted_fields = (sysv, roots_all, bounds)
        session.late_by_system = {}
        if not special and not may_merge:
            return (session.results() if need_records else None), keep, {}, set()
        known = None
        if not need_records and hasattr(session, "sparse_records"):
            # only the systems with a collector or with two candidates for a merge need their full records
            need = set(np.unique(sysv[is_coll | is_member]).tolist())
            if may_merge:
                u, cnt = np.unique(touch_sys, return_counts=True)
                need |= set(u[cnt >= 2].tolist())
            hint = int(sum(bounds[s + 1] - bounds[s] for s in need))
            if session.sparse_records(need, hint):
                known = (np.sort(slim["index"][is_coll].astype(np.int64)), np.sort(slim["index"][is_member].astype(np.int64)))
        res = session.results()
        if special:
            alive, dup, extra = resolve_collectors(res, known=known)
        else:
            alive, dup, extra = np.ones(len(res), dtype=bool), {}, set()
        n0 = len(res)
        if merge:
            cand = slim["index"][touch_pos].astype(np.int64)
            if known is not None:                              # sparse tables: lone candidates of other systems read as zeros
                cand = cand[np.isin(touch_sys, np.fromiter(need, dtype=np.int64, count=len(need)))]
            res, alive, dup, extra = merge_exterior(session, res, alive, dup, extra, cand=cand[alive[cand]])
        dead = np.nonzero(~alive[:n0])[0]
        late = np.nonzero(alive[n0:])[0] + n0
        by_sys = {}
        for i in late.tolist():
            by_sys.setdefault(int(res["system"][i]), []).append(i)
        session.late_by_system = by_sys
        session.redo_systems = set(np.unique(res["system"][dead]).tolist()) | set(by_sys)
        session.alive = alive
        touched = getattr(session, "patched_roots", None)
        if touched:                                             # roots rewritten by a merge after the device gathered them
            session.redo_systems |= {int(res["system"][i]) for i in touched}
        return res, keep, dup, extra
    res = session.results()
    alive, dup, extra = resolve_collectors(res)
    if merge:
        res, alive, dup, extra = merge_exterior(session, res, alive, dup, extra)
    order, base, n_sorted = session.dfs_order() if hasattr(session, "dfs_order") else (None, 0, 0)
    if order is not None and n_sorted:
        # records [base, base + n_sorted) arrive with their reporting order from the device (radix sort); the others
        # (first levels of a large system, re-solved merged boxes: a handful) are sorted here and merged in
        mask = alive[base:base + n_sorted][order["index"]]
        packed = order[mask]                                    # (index, system, root) of the survivors, in reporting order
        main = packed["index"].astype(np.int64) + base
        late = np.r_[np.nonzero(alive[:base])[0], np.nonzero(alive[base + n_sorted:])[0] + base + n_sorted]
        # late records stay out of `keep`; their systems are assembled individually in per_system_roots
        by_sys = {}
        for i in late.tolist():
            by_sys.setdefault(int(res["system"][i]), []).append(i)
        session.late_by_system = by_sys
        keep = main
        session.sorted_fields = (packed["system"], packed["root"])
        touched = getattr(session, "patched_roots", None)
        if touched:                                             # roots rewritten by a merge after the device gathered them
            sysv, roots = session.sorted_fields
            roots = np.array(roots)
            for i in touched:
                roots[keep == i] = res["root"][i]
            session.sorted_fields = (sysv, roots)
    else:
        keep = _path_order(res, np.nonzero(alive)[0])
        session.sorted_fields = None
        session.late_by_system = {}
    return res, keep, dup, extra


def per_system_roots(res, keep, dup, extra, nsys, dim, want_boxes=False, sorted_fields=None, late_by_system=None,
                     redo_systems=None, alive=None):
    """Vectorised gathering of roots per system; special records are patched in afterwards.  `sorted_fields` =
    (system, root) of the kept records in order, when the device already gathered them; `late_by_system` = live
    records that are not in `keep` (produced after the device sorted), merged into their systems by path;
    `redo_systems` = systems whose slice of `keep` contains dead records (`alive` says which)."""
    bounds = None
    if sorted_fields is not None:
        sysv, roots_all = sorted_fields[:2]
        bounds = sorted_fields[2] if len(sorted_fields) > 2 else None
    else:
        sysv = np.take(res["system"], keep)
        roots_all = np.take(res["root"], keep, axis=0)
    if bounds is None:
        bounds = np.searchsorted(sysv, np.arange(nsys + 1))
    late_by_system = late_by_system or {}
    redo_systems = redo_systems or set()
    out_roots, out_boxes, has_dup, has_extra = [], [], False, False
    special = {int(k) for k in dup} | {int(k) for k in extra}
    special_systems = {int(res["system"][k]) for k in special} if special else set()
    if not want_boxes:
        # ordinary systems are slices of the sorted roots: one split for all of them, then only the special ones
        # (lost records to a merge / collector, late records, duplicate or extra roots) are rebuilt record by record
        bl = bounds.tolist()
        out_roots = [roots_all[bl[s]:bl[s + 1]] for s in range(nsys)]
        todo = sorted((set(redo_systems) | set(late_by_system) | special_systems) & set(range(nsys)))
    else:
        todo = range(nsys)
    for s in todo:
        idx = keep[bounds[s]:bounds[s + 1]]
        redo = s in redo_systems
        if redo:
            idx = idx[alive[idx]]
        late = late_by_system.get(s)
        if late:
            idx = _path_order(res, np.r_[idx.astype(np.int64), np.asarray(late, dtype=np.int64)])
        if late or redo or s in special_systems:
            rows = []
            for i in idx:
                if int(i) in dup:
                    rows += dup[int(i)]
                    has_dup = True
                else:
                    rows.append(res["root"][i])
                if int(i) in extra:
                    has_extra = True
            r = np.array(rows).reshape(-1, dim)
        else:
            r = roots_all[bounds[s]:bounds[s + 1]]
        if want_boxes:
            out_roots.append(r)
            out_boxes.append([ResultBox(res[i], dup.get(int(i)), int(i) in extra) for i in idx])
        else:
            out_roots[s] = r
    if has_extra:
        warnings.warn("Might Have Extra Roots! See Bounding Boxes for details!")
    if has_dup:
        warnings.warn("Might Have Duplicate Roots! See Bounding Boxes for details!")
    return out_roots, out_boxes
