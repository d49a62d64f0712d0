"""CPU, world_size 2 (gloo): the N>1 plumbing of the object-sharded mapping path."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dqo_map_b200 import sharding


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        counts = {0: 700, 1: 300, 2: 250, 3: 50, 4: 20}
        owner, load = sharding.assign_objects(counts, world)
        mine = sharding.local_objects(owner, rank)
        table = torch.tensor([[float(o), float(counts[o])] + [float(o) * 0.5] * 10 for o in mine], dtype=torch.float32)
        table = table.reshape(len(mine), 12)
        full = sharding.gather_object_table(table)
        assert full.shape == (5, 12)
        assert full[:, 0].tolist() == [0.0, 1.0, 2.0, 3.0, 4.0]
        assert full[:, 1].tolist() == [700.0, 300.0, 250.0, 50.0, 20.0]
        fixed = sharding.gather_object_table(torch.full((2, 12), float(rank)), rows_per_rank=2)
        assert fixed.shape == (4, 12) and fixed[:, 0].tolist() == [0.0, 0.0, 1.0, 1.0]
        n_local = sum(counts[o] for o in mine)
        g = {"xyz": torch.full((n_local, 3), float(rank)), "opacity": torch.full((n_local, 1), 0.5 + rank)}
        got = sharding.gather_gaussians(g, dst=0)
        if rank == 0:
            assert got["xyz"].shape == (sum(counts.values()), 3)
            assert int((got["xyz"][:, 0] == 0).sum()) == load[0] and int((got["xyz"][:, 0] == 1).sum()) == load[1]
            assert got["opacity"].shape == (sum(counts.values()), 1)
        else:
            assert got is None
        results[rank] = True
    finally:
        dist.destroy_process_group()


def test_gather_paths_world_size_2():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert dict(results) == {0: True, 1: True}
