"""world_size-2 gloo tests of the N>1 host logic (no GPU): sharding, weight broadcast, result gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from premvos_b200 import shard


def test_shard_units_partition():
    for n in (0, 1, 7, 89, 90):
        for world in (1, 2, 4, 8):
            parts = [shard.shard_units(n, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard.shard_units(4, 2, 2)
    assert shard.pairs_of_video(["a", "b", "c"]) == [("a", "b"), ("b", "c")]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sd = {"w": torch.arange(6, dtype=torch.float32).reshape(2, 3), "b": torch.ones(3)} if rank == 0 else {}
        sd = shard.broadcast_state_dict(sd, src=0)
        assert torch.equal(sd["w"], torch.arange(6, dtype=torch.float32).reshape(2, 3)) and torch.equal(sd["b"], torch.ones(3))
        mine = shard.shard_units(5, rank, world)
        local = {i: np.full((2, 2), float(i), dtype=np.float32) * sd["b"][0].item() for i in mine}
        merged = shard.gather_results(local, dst=0)
        t = torch.tensor([float(len(mine))])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)       # the bench's max-over-ranks step
        if rank == 0:
            assert sorted(merged) == [0, 1, 2, 3, 4] and all(float(merged[i][0, 0]) == i for i in merged)
            q.put(("ok", float(t.item())))
        else:
            assert merged is None
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_roundtrip():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == ("ok", 3.0)
