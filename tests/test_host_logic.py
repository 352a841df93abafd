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


def test_window_index_table_matches_per_video_gather():
    """The multi-video window table addresses the concatenated frame axis exactly as gather_windows does per video
    (incl. the repeat-last-frame rule for a video shorter than the window)."""
    lengths = [7, 20, 33, 21, 1]
    feats = [torch.randn(t, 3, generator=torch.Generator().manual_seed(t)) for t in lengths]
    starts, idx = windowing.window_index_table(lengths, 20, 8)
    assert starts == [windowing.window_starts(t, 20, 8) for t in lengths]
    got = torch.cat(feats)[idx]
    want = torch.cat([windowing.gather_windows(f, st, 20) for f, st in zip(feats, starts)])
    assert torch.equal(got, want)


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


def test_bench_rank0_only_section_has_no_collective():
    """bench.py's records after `if rank != 0: return` run on rank 0 alone while the other ranks wait at the
    final barrier: a collective there (e.g. a trainer step that all-reduces its gradients) deadlocks the
    multi-GPU launch.  Source-level guard over the functions that section calls."""
    import ast
    import inspect
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "bench.py")).read()
    tree = ast.parse(src)
    funcs = {n.name: n for n in tree.body if isinstance(n, ast.FunctionDef)}
    run_infer = funcs["run_infer"]
    cut = next(i for i, st in enumerate(run_infer.body)
               if isinstance(st, ast.If) and "rank != 0" in ast.get_source_segment(src, st.test))
    called = set()
    for st in run_infer.body[cut + 1:]:
        for n in ast.walk(st):
            if isinstance(n, ast.Call) and isinstance(n.func, ast.Name) and n.func.id in funcs:
                called.add(n.func.id)
            if isinstance(n, ast.Name) and n.id in funcs:            # functions passed by name (side-record table)
                called.add(n.id)
    assert {"measure_head_only", "measure_alt_heads", "hbm_rooflines", "time_dominant_conv"} <= called
    for name in sorted(called):
        body = ast.get_source_segment(src, funcs[name])
        assert "dist." not in body and "_timed(" not in body, f"{name} touches the process group"
        for n in ast.walk(funcs[name]):
            if isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute) and n.func.attr == "step":
                kw = {k.arg: k.value for k in n.keywords}
                assert isinstance(kw.get("sync_grads"), ast.Constant) and kw["sync_grads"].value is False, \
                    f"{name}: trainer.step() on rank 0 alone must pass sync_grads=False"
    # and the switch exists on both trainers
    from feature_vs_text_compound_emotion_b200 import heads_training, training
    for cls in (training.HeadTrainer, heads_training.AltHeadTrainer):
        assert "sync_grads" in inspect.signature(cls.step).parameters


def test_recorded_bench_line_carries_the_contract_keys():
    """The last bench line recorded on a B200 (profiles/r02o_bench_n1.json, printed by `python bench.py`) has every
    key of the bench contract, with consistent values: value = frames / time, roofline.frac = achieved / peak,
    e2e measured with host copies, launches counted, the CPU baseline run on the reference's own modules."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    line = json.loads(open(os.path.join(root, "profiles", "r02o_bench_n1.json")).read())
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["metric"] == "frames_per_s" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["n_gpus"] == 1 and line["warmup"] >= 3 and line["data"] == "synthetic" and "workload" in line["config"]
    frames = line["config"]["frames_per_step_per_gpu"]
    assert abs(line["value"] - frames / (line["ms_per_step"] * 1e-3)) <= 1e-6 * line["value"]
    e2e = line["e2e"]
    assert e2e["unit"] == "frames/s" and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0
    assert 0 < e2e["value"] <= 1.02 * line["value"]                 # copies cannot make it faster (2 % clock noise)
    r = line["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0.5 < r["frac"] < 1.0 and r["kernel"].startswith(r["plan_variant_for_this_layer"])
    assert r["traffic"] is None or r["traffic"] > 0
    c = line["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and c["unit"] == "frames/s" and c["value"] > 0 and c["sample"]
    assert line["gpu_launches"] > 0 and line["gpu_launches"] % line["steps"] == 0
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    for sub in ("long_run", "ir50", "ir50_layers", "roofline_hbm", "head_only", "train", "full", "sweep", "alt_heads"):
        assert sub in line, sub
