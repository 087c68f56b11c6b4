"""CPU, world_size 2, gloo: the mesh-sharding plan and the statistics reduction used by bench.py
for N > 1 (no data-path collective exists on this path)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from harry_b200 import shard


def test_plan_round_robin_and_balanced():
    assert shard.shard_plan(5, 2) == [[0, 2, 4], [1, 3]]
    plan = shard.shard_plan(6, 3, sizes=[10, 1, 1, 9, 8, 2])
    assert sorted(sum(plan, [])) == list(range(6))
    loads = [sum([10, 1, 1, 9, 8, 2][i] for i in p) for p in plan]
    assert max(loads) - min(loads) <= 2


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.shard_plan(7, world)[rank]
    units = float(sum(100 * (i + 1) for i in mine))
    t, u = shard.reduce_stats(dist, torch.device("cpu"), 10.0 * (rank + 1), units)
    if rank == 0:
        out.put((t, u))
    dist.barrier()
    dist.destroy_process_group()


def test_reduce_stats_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    t, u = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert t == 20.0
    assert u == float(sum(100 * (i + 1) for i in range(7)))


def test_gpu_local_cpus(tmp_path):
    """bind_rank_to_gpu_node reads the GPU's local_cpulist from sysfs: list syntax, missing entries"""
    from harry_b200 import shard
    assert shard._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert shard._parse_cpulist("\n") == set()
    dev = tmp_path / "0000:e5:00.0"
    dev.mkdir()
    (dev / "local_cpulist").write_text("0-1,64-65\n")
    assert shard.gpu_local_cpus(0, 0xE5, 0, str(tmp_path)) == {0, 1, 64, 65}
    assert shard.gpu_local_cpus(0, 0xE6, 0, str(tmp_path)) == set()
