"""Multi-GPU: the batch shards by system, one process per GPU, no data-path collective.

Systems of a batch never interact, so rank r solves systems r, r+W, r+2W, ... on its own
device frontier.  The only exchange is at the end: an all-gather of the per-system root counts
followed by an all-gather of the (padded) roots, so that every rank returns the full answer --
NCCL over NVLink on GPUs, gloo in the CPU tests.  No floating-point reduction is involved, so the
result does not depend on the number of ranks.
"""
import numpy as np


def shard(n_items, rank, world):
    """Indices of the systems rank `rank` owns (round-robin)."""
    return np.arange(rank, n_items, world)


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
        host = torch.empty(big.shape, dtype=torch.float64, pin_memory=True)
        host.copy_(big, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        blocks = host.numpy()
    else:
        blocks = big.numpy()
    out = [None] * n_systems
    for r in range(world):
        ids = shard(n_systems, r, world)
        if not len(ids):
            continue
        ends = np.cumsum(counts_h[ids])
        for s_id, piece in zip(ids.tolist(), np.split(blocks[r][:int(ends[-1])], ends[:-1])):
            out[s_id] = piece
    return out


def solve_batch_sharded(systems, errors, group=None, engine=None, **kw):
    """`solve_batch` over all ranks of the process group; every rank gets all roots."""
    import torch.distributed as dist
    from .ChebyshevSubdivisionSolver import solve_batch
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ids = shard(len(systems), rank, world)
    local = solve_batch([systems[i] for i in ids], [errors[i] for i in ids], engine=engine, **kw) if len(ids) else []
    return all_gather_roots(local, ids, len(systems), len(systems[0]), group=group)
