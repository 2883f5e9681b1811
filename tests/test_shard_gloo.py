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
        # the data-plane collective of the sharded path: one flat uint8 buffer per rank and step, gathered as a tensor (no pickle),
        # asynchronously (two buffers in flight, as bench.py does it)
        works = []
        for step in range(3):
            flat = torch.full((1000,), 10 * step + rank, dtype=torch.uint8)
            bucket, work = shard.gather_tensor(flat, dst=0, async_op=True)
            works.append((step, bucket, work))
        for step, bucket, work in works:
            work.wait()
            if rank == 0:
                assert len(bucket) == world and all(int(bucket[r][0]) == 10 * step + r and int(bucket[r][-1]) == 10 * step + r for r in range(world))
            else:
                assert bucket is None
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


def test_combine_proposals_matches_the_json_round_trip():
    """pipeline.combine_proposals == detect_one_image post-processing (eval.py:93-96) + convert_results_to_json
    (train.py:399-408) + combine_general_and_specific.py:37 on the same boxes."""
    import json
    from premvos_b200 import pipeline, propnet
    rng = np.random.default_rng(0)
    H, W = 480, 854
    Hp, Wp = propnet.custom_resize_shape(H, W)
    assert (Hp, Wp) == (749, 1333) and pipeline.flow_input_shape(436, 1024) == (448, 1024) and pipeline.flow_input_shape(480, 854) == (512, 896)
    sets = []
    for n in (5, 0, 20):
        x1 = rng.uniform(-20, Wp - 50, n); y1 = rng.uniform(-20, Hp - 50, n)
        sets.append(np.stack([x1, y1, x1 + rng.uniform(5, 400, n), y1 + rng.uniform(5, 400, n)], 1).astype(np.float32).reshape(-1, 4))
    for g, s in ((sets[0], sets[2]), (sets[1], sets[0]), (sets[1], sets[1])):
        ref = []
        for boxes in (g, s):
            scale = (Hp * 1.0 / H + Wp * 1.0 / W) / 2
            b = propnet.clip_boxes(boxes.copy() / scale, (H, W))
            res = [propnet.SecondDetectionResult(bb, 0.9, 1, None, None, 1, None, None) for bb in b]
            ref += json.loads(json.dumps(propnet.convert_results_to_json(res)))
        got = pipeline.combine_proposals(g, s, (H, W), (Hp, Wp))
        assert got.shape == (len(ref), 4) and got.dtype == np.float32
        if ref:
            np.testing.assert_array_equal(got, np.array([r["bbox"] for r in ref], dtype=np.float32))
