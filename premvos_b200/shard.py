"""Frame-pair sharding across ranks.

The units of this path are independent (stage 1: frame pairs, script_pwc_multi.py:100; the
reference itself only offers manual job slicing, MergeTrack/merge.py:126-129), so ranks never
exchange activations: rank r owns units r, r+R, r+2R, ...  The only collectives are the one-time
weight broadcast and the gather of per-unit results to rank 0."""
from __future__ import annotations

import os
from typing import List, Sequence


def env_rank_world():
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when absent."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_units(num_units: int, rank: int, world: int) -> List[int]:
    """Indices of the units rank `rank` of `world` processes: round-robin, so consecutive frames
    (similar cost) spread evenly and the tail imbalance is at most one unit."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %d/%d" % (rank, world))
    return list(range(rank, num_units, world))


def pairs_of_video(frames: Sequence) -> list:
    """Consecutive frame pairs (t, t+1) of one video, script_pwc_multi.py:100."""
    return list(zip(frames[:-1], frames[1:]))


def broadcast_state_dict(sd: dict, src: int = 0):
    """Broadcast a {name: tensor} dict from `src` (weights are loaded once, on rank 0)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return sd
    names = [None]
    if dist.get_rank() == src:
        names = [[(k, tuple(v.shape)) for k, v in sd.items()]]
    dist.broadcast_object_list(names, src=src)
    out = {}
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    for k, shp in names[0]:
        t = (sd[k].to(torch.float32).to(dev).contiguous() if dist.get_rank() == src
             else torch.empty(shp, dtype=torch.float32, device=dev))
        dist.broadcast(t, src=src)
        out[k] = t.cpu()
    return out


def gather_results(local: dict, dst: int = 0):
    """Gather {unit index: numpy array} dicts to rank `dst` (returns the merged dict there, None elsewhere)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(local)
    bucket = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(local, bucket, dst=dst)
    if dist.get_rank() != dst:
        return None
    merged = {}
    for d in bucket:
        merged.update(d)
    return merged


def gather_tensor(t, dst: int = 0, async_op: bool = False):
    """One equally-shaped tensor per rank -> list of `world` tensors on rank `dst` (None elsewhere), as ONE collective on the
    tensors' own device: NCCL over NVLink for CUDA tensors (the packed per-unit results of a step, pipeline.pack_step_results;
    SURVEY.md section 8e), gloo for CPU tensors.  No pickling, no host round trip.  async_op=True returns (bucket, work): the
    collective runs on the backend's stream, call work.wait() before reading `bucket` or overwriting `t`."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return ([t], None) if async_op else [t]
    bucket = [torch.empty_like(t) for _ in range(dist.get_world_size())] if dist.get_rank() == dst else None
    work = dist.gather(t, bucket, dst=dst, async_op=async_op)
    return (bucket, work) if async_op else bucket
