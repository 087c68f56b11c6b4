"""Sharding of independent meshes over ranks: the attribute path has no exchange
step, so a batch is partitioned by mesh and nothing but the final timing / size statistics is
combined.  Works with any torch.distributed backend (NCCL on the B200 box, gloo in CPU tests)."""
from __future__ import annotations


def shard_plan(n_meshes: int, world: int, sizes=None) -> list:
    """Mesh ids per rank.  Without sizes: round robin (mesh i -> rank i mod world).  With sizes:
    greedy longest-processing-time balancing, deterministic."""
    plan = [[] for _ in range(world)]
    if sizes is None:
        for i in range(n_meshes):
            plan[i % world].append(i)
        return plan
    load = [0] * world
    for i in sorted(range(n_meshes), key=lambda k: (-sizes[k], k)):
        r = min(range(world), key=lambda k: (load[k], k))
        plan[r].append(i)
        load[r] += sizes[i]
    for p in plan:
        p.sort()
    return plan


def reduce_stats(dist, device, elapsed_ms: float, units: float):
    """max over ranks of the elapsed time, sum over ranks of the processed units."""
    import torch
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    u = torch.tensor([units], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())
