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


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_local_cpus(pci_domain: int, pci_bus: int, pci_device: int, sysfs: str = "/sys/bus/pci/devices") -> set:
    """CPUs on the NUMA node the GPU hangs off (sysfs local_cpulist); empty when the kernel does not say."""
    path = f"{sysfs}/{pci_domain:04x}:{pci_bus:02x}:{pci_device:02x}.0/local_cpulist"
    try:
        with open(path) as f:
            return _parse_cpulist(f.read())
    except (OSError, ValueError):
        return set()


def bind_rank_to_gpu_node(device: int, sysfs: str = "/sys/bus/pci/devices"):
    """One process per GPU: keep the rank on the CPUs next to its GPU, so that the host buffers it allocates afterwards
    (first touch) and the staging copies stay on the socket whose PCIe root the GPU is attached to -- on a two-socket
    box the ranks of the far GPUs otherwise push every upload through the inter-socket link.  Returns the CPU set the
    process runs on afterwards (unchanged when the topology is flat, unknown, or outside the allowed set)."""
    import os
    allowed = os.sched_getaffinity(0)
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        local = gpu_local_cpus(int(p.pci_domain_id), int(p.pci_bus_id), int(p.pci_device_id), sysfs)
    except Exception:
        return allowed
    want = allowed & local
    if want and want != allowed:
        os.sched_setaffinity(0, want)
        return want
    return allowed
