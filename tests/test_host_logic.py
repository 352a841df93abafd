"""Host-side logic on the CPU: window grid vs the reference fixture, window gathering, video
sharding, and the world-size-2 gather over gloo."""
import json
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from feature_vs_text_compound_emotion_b200 import sharding, windowing


def test_window_starts_match_reference_fixture(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "windowing.json")))
    for L, wins in cases.items():
        L = int(L)
        starts = windowing.window_starts(L, 300, 200)
        assert starts == [w[0] for w in wins], L
        if L >= 300:
            assert all(w[2] == 300 for w in wins) and starts[-1] + 300 == L or (L - 300) % 200 == 0


def test_gather_windows_and_short_video_padding():
    feat = torch.arange(701, dtype=torch.float32).view(-1, 1)
    w = windowing.gather_windows(feat, windowing.window_starts(701), 300)
    assert w.shape == (4, 300, 1) and w[3, 0, 0] == 401 and w[1, 0, 0] == 200
    short = torch.arange(5, dtype=torch.float32).view(-1, 1)
    w = windowing.gather_windows(short, [0], 8)
    assert w[0, :, 0].tolist() == [0, 1, 2, 3, 4, 4, 4, 4]      # last frame repeated


def test_shard_videos_balanced_and_complete():
    g = torch.Generator().manual_seed(7)
    lengths = torch.randint(150, 3001, (56,), generator=g).tolist()      # C-EXPR-DB-CHALLENGE has 56 videos
    for world in (1, 2, 4, 8):
        shards = sharding.shard_videos(lengths, world)
        assert sorted(i for s in shards for i in s) == list(range(56))
        loads = [sum(lengths[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(lengths)


def _worker(rank, world, port, lengths, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def infer_one(i):        # stands in for the GPU pipeline: logits that identify (video, frame)
            t = torch.arange(lengths[i], dtype=torch.float32).view(-1, 1)
            return torch.cat([t, torch.full_like(t, float(i))], dim=1)
        outs = sharding.run_sharded(infer_one, lengths, 2, torch.device("cpu"))
        ok = all(o.shape == (lengths[i], 2) and bool((o[:, 1] == i).all()) and
                 bool((o[:, 0] == torch.arange(lengths[i])).all()) for i, o in enumerate(outs))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_sharded_gather_world_size_2_gloo():
    lengths = [310, 150, 977, 300, 45, 620, 1201]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def _ar_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from feature_vs_text_compound_emotion_b200.training import all_reduce_flat
        flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        scale = all_reduce_flat(flat)
        q.put((rank, scale, bool(torch.equal(flat, torch.arange(1000, dtype=torch.float32) * 3))))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_bucket_all_reduce_world_size_2_gloo():
    """The training step's only collective: SUM over the flat bucket, 1/world returned as the scale."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_ar_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, 0.5, True), (1, 0.5, True)]
    from feature_vs_text_compound_emotion_b200.training import all_reduce_flat
    assert all_reduce_flat(torch.ones(4)) == 1.0          # no process group: nothing to do
