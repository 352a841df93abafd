"""Video-level sharding across the GPUs of one box.

Every video is independent (eval-mode BN, windows recomputed from scratch), so ranks never talk
on the data path: each rank runs the full pipeline on its shard and one all_gather of the
per-frame logits closes the job.  Works with any torch.distributed backend (nccl on GPUs, gloo in
the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence

import torch


def shard_videos(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy longest-first assignment: videos sorted by decreasing length, each given to the
    currently lightest rank.  Returns the video indices of every rank."""
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    load = [0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += lengths[i]
    return shards


def gather_predictions(local: Dict[int, torch.Tensor], lengths: Sequence[int], n_out: int, device) -> List[torch.Tensor]:
    """All ranks end up with the per-frame logits of every video.  `local` maps video index ->
    [T_v, n_out] for the videos this rank ran.  One all_gather on a frame-padded buffer."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    shards = shard_videos(lengths, world)
    cap = max(sum(lengths[i] for i in s) for s in shards) if lengths else 0
    rank = dist.get_rank() if dist.is_initialized() else 0
    buf = torch.zeros(max(cap, 1), n_out, dtype=torch.float32, device=device)
    off = 0
    for i in shards[rank]:
        buf[off:off + lengths[i]] = local[i]
        off += lengths[i]
    if world > 1:
        allbuf = torch.empty(world * buf.shape[0], n_out, dtype=torch.float32, device=device)
        dist.all_gather_into_tensor(allbuf, buf)
        allbuf = allbuf.view(world, buf.shape[0], n_out)
    else:
        allbuf = buf.unsqueeze(0)
    out: List[torch.Tensor] = [None] * len(lengths)
    for r, s in enumerate(shards):
        off = 0
        for i in s:
            out[i] = allbuf[r, off:off + lengths[i]]
            off += lengths[i]
    return out


def run_sharded(infer_one: Callable[[int], torch.Tensor], lengths: Sequence[int], n_out: int, device) -> List[torch.Tensor]:
    """infer_one(video_index) -> [T_v, n_out] on `device`.  Each rank runs its shard, then all ranks
    gather everything."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    mine = shard_videos(lengths, world)[rank]
    local = {i: infer_one(i) for i in mine}
    return gather_predictions(local, lengths, n_out, device)
