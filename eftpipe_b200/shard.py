"""Batch sharding over the GPUs of one node.

The hot path has no exchange step: every point of the batch is independent and the plan constants are
read-only, so the batch axis is split into contiguous shards, one process per GPU evaluates its shard with its
own replica of the plan, and the only collective is the final gather of the per-point results (`logp`,
`status`, optional best-fit vectors) - 8 bytes per point.  The reference has no counterpart (its only
parallelism is Cobaya's independent MPI chains); this module is the host-side plumbing
for samplers that hand a whole population of points to one call.

Works with any `torch.distributed` backend: `nccl` on the GPUs, `gloo` in the CPU tests
(tests/test_sharding.py, world_size 2).
"""
from __future__ import annotations


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous [lo, hi) of `rank` when `n` points are split over `world` ranks; the first `n % world` ranks
    get one extra point (numpy.array_split convention)."""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError(f"invalid shard request n={n} world={world} rank={rank}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_points(local, n_total: int, group=None):
    """All-gather per-point results.  `local`: tensor (n_local, ...) of this rank's shard (shard_bounds order);
    returns the (n_total, ...) tensor on every rank.  Ragged shards are padded to the largest one for the
    collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if local.shape[0] != n_total:
            raise ValueError("single-process gather: shard does not cover the batch")
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(n_total, world, rank)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank}: shard has {local.shape[0]} points, expected {hi - lo}")
    width = -(-n_total // world)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: hi - lo] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = []
    for r, p in enumerate(parts):
        a, b = shard_bounds(n_total, world, r)
        out.append(p[: b - a])
    return torch.cat(out, dim=0)


def evaluate_sharded(local_eval, n_total: int, group=None):
    """Run `local_eval(lo, hi)` (returns a tensor with leading axis hi - lo, or a tuple of such tensors) on
    this rank's shard and gather the results of all ranks."""
    import torch.distributed as dist

    world, rank = 1, 0
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(n_total, world, rank)
    res = local_eval(lo, hi)
    if isinstance(res, tuple):
        return tuple(gather_points(r, n_total, group) for r in res)
    return gather_points(res, n_total, group)
