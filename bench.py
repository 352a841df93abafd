#!/usr/bin/env python
"""bench.py -- end-to-end frames/s of the LFAN inference hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference arithmetic on host cores

Workload (config.workload): BASELINE.json configs[1] shape -- 8 windows x 300 aligned 40x40 face
crops per GPU -- run through the WHOLE path the metric names (IR-50 -> TCN x3 -> cross-modal
attention -> classifier) with synthetic VGGish/BERT feature windows beside the frames.
One step = one forward over that batch (2400 frames per GPU).  Weak scaling: every rank runs
its own 8 windows; the only collective is the final all_gather of per-frame logits.

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM (CUDA events, max
over ranks); `e2e` = the same through the public nn.Module API with pinned-host inputs, H2D and
D2H inside the timed region; `roofline` = the dominant kernel (tcgen05 implicit-GEMM conv,
256->256 @10x10 class = 50.7 % of the FLOPs; CTA-pair cta_group::2 variant) timed alone; `cpu_baseline` = the oracle restatement
of the reference on this box's host cores over a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

import torch  # noqa: E402

WINDOWS, LENGTH = 8, 300
MODS = ["video", "vggish", "bert"]
IR50_GFLOP_PER_FRAME = 6.0545            # SURVEY.md section 8d / BASELINE.md section 2
HEAD_MFLOP_PER_FRAME = 9.988
DOM = dict(h=10, cin=256, cout=256, ksize=3)   # dominant conv class
BS = {"visual_state_dict": "res50_ir_0.887", "audio_state_dict": "vggish"}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:  # NVML missing: report that instead of inventing numbers
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def build_model(dev):
    from feature_vs_text_compound_emotion_b200 import synthetic
    from feature_vs_text_compound_emotion_b200.models.model import LFAN
    m = LFAN(backbone_settings=BS, output_dim=7, task="CLASSIFICATION", modality=MODS, kernel_size=5,
             example_length=LENGTH, tcn_channel=synthetic.TCN_CHANNELS, modal_dim=32, num_heads=2, root_dir="", device=dev)
    m.init(visual_state_dict=synthetic.visual_backbone_state_dict(0))
    m.load_state_dict(synthetic.lfan_state_dict(0, MODS), strict=True)
    return m.to(dev).eval()


def host_batch(seed):
    from feature_vs_text_compound_emotion_b200 import synthetic
    f = synthetic.feature_windows(WINDOWS, LENGTH, seed=seed, modalities=["vggish", "bert"])
    return {"video": synthetic.frames(WINDOWS * LENGTH, seed=seed + 1).view(WINDOWS, LENGTH, 3, 40, 40),
            "vggish": f["vggish"], "bert": f["bert"]}


def time_dominant_conv(dev, frames, iters=20):
    """The 256->256 3x3 @10x10 conv (PReLU epilogue), alone, back to back: average launch time."""
    from feature_vs_text_compound_emotion_b200.engine import conv_forward
    g = torch.Generator().manual_seed(0)
    h, cin, cout = DOM["h"], DOM["cin"], DOM["cout"]
    x = torch.randn(frames + 8, h, h, cin, generator=g).to(torch.bfloat16).to(dev)
    w = (torch.randn(cout, 9 * cin, generator=g) * (9 * cin) ** -0.5).to(torch.bfloat16).to(dev)
    bias = torch.randn(9, cout, generator=g).to(dev)
    alpha = torch.full((cout,), 0.25).to(dev)
    for _ in range(3):
        conv_forward(x, w, bias, 3, 1, 1, alpha=alpha, n_frames=frames)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        conv_forward(x, w, bias, 3, 1, 1, alpha=alpha, n_frames=frames)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * frames * h * h * cout * 9 * cin
    return ms, flops


def cpu_reference_step(sd, frames, threads):
    """One bounded step of the reference arithmetic (oracle restatement) on host cores:
    `frames` frames through IR-50 plus one 300-frame head window."""
    from feature_vs_text_compound_emotion_b200 import synthetic
    from oracle import lfan_oracle as O
    torch.set_num_threads(threads)
    x = synthetic.frames(frames, seed=5)
    feats = synthetic.feature_windows(1, LENGTH, seed=6, modalities=["vggish", "bert"])
    t0 = time.perf_counter()
    with torch.no_grad():
        emb = O.ir50_forward(sd, x, "spatial.visual.backbone.")
        reps = (LENGTH + frames - 1) // frames
        vid = emb.repeat(reps, 1)[:LENGTH].view(1, LENGTH, -1)
        O.head_forward(sd, {"video": vid, "vggish": feats["vggish"].squeeze(1), "bert": feats["bert"].squeeze(1)}, MODS)
    dt = time.perf_counter() - t0
    # the head ran on a full 300-frame window; charge it pro rata to the `frames` sample
    return dt


def run_reference(args):
    """--impl reference: the reference's CPU arithmetic (oracle port; the Python reference itself
    cannot travel to the GPU box) on all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from feature_vs_text_compound_emotion_b200 import synthetic
    threads = os.cpu_count() or 1
    sd = synthetic.lfan_state_dict(0, MODS)
    t_cal = cpu_reference_step(sd, 16, threads)
    budget = 150.0 / max(1, args.steps + args.warmup)
    frames = int(max(16, min(LENGTH, 16 * budget / max(t_cal, 1e-3))))
    for _ in range(args.warmup):
        cpu_reference_step(sd, frames, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(sd, frames, threads)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    fps = frames / dt
    sample = f"{frames} frames through IR-50 + one {LENGTH}-frame head window per step (oracle port, fp32, torch CPU)"
    print(json.dumps({
        "impl": "reference", "metric": "frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"LFAN inference (IR-50+TCN+fusion), {WINDOWS} windows x {LENGTH} frames of 40x40 crops per GPU",
                   "sample": sample},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def _timed(fn, steps, warmup, world, dist, dev, local):
    """warm-up, barrier, CUDA-event timing of `steps` calls, max over ranks -> (total ms, clocks)."""
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item(), clocks


def video_lengths(n, seed=7):
    """SURVEY.md section 8d cfg 3/5: T_v ~ U{150..3000}, seeded."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(150, 3001, (n,), generator=g).tolist()


def run_videos(args, dev, world, rank, local, dist):
    """configs[2] / configs[4]: whole videos from stored uint8 256x256 crops + per-frame log-mel
    examples + BERT features -> device eval transform -> IR-50 / VGGish -> TCN -> fusion -> window
    stitching -> video-level vote.  Videos are sharded over the ranks (longest first); the only
    collective is the final gather of per-frame logits."""
    from feature_vs_text_compound_emotion_b200 import sharding, synthetic, windowing
    from feature_vs_text_compound_emotion_b200.models.model import LFAN
    mods = ["video", "logmel", "bert"]
    m = LFAN(backbone_settings=BS, output_dim=7, task="CLASSIFICATION", modality=mods, kernel_size=5, example_length=LENGTH,
             tcn_channel=synthetic.TCN_CHANNELS, modal_dim=32, num_heads=2, root_dir="", device=dev)
    m.init(visual_state_dict=synthetic.visual_backbone_state_dict(0), audio_state_dict=synthetic.vggish_state_dict(0))
    m.load_state_dict(synthetic.lfan_state_dict(0, mods), strict=True)
    m = m.to(dev).eval()
    n_videos = args.videos or (56 if args.workload == "full" else 1000)
    lengths = video_lengths(n_videos)
    mine = sharding.shard_videos(lengths, world)[rank]
    tmax = max(lengths)
    # one synthetic video of the maximum length; every video is a prefix of it (synthetic data, real sizes)
    raw_host = synthetic.raw_frames_u8(64, seed=11 + rank).repeat((tmax + 63) // 64, 1, 1, 1)[:tmax].contiguous().pin_memory()
    lm_host = synthetic.logmel_patches(tmax, seed=12 + rank).pin_memory()
    bert_host = torch.randn(tmax, 768, generator=torch.Generator().manual_seed(13 + rank)).pin_memory()
    raw, lm, bert = raw_host.to(dev), lm_host.to(dev), bert_host.to(dev)
    votes = {}

    def run_shard(from_host):
        local_out = {}
        for i in mine:
            T = lengths[i]
            if from_host:
                r, a, b = raw_host[:T].to(dev, non_blocking=True), lm_host[:T].to(dev, non_blocking=True), bert_host[:T].to(dev, non_blocking=True)
            else:
                r, a, b = raw[:T], lm[:T], bert[:T]
            local_out[i] = windowing.infer_video(m, r, {"logmel": a, "bert": b})
            votes[i] = local_out[i]
        return sharding.gather_predictions(local_out, lengths, 7, dev)

    total_ms, clocks = _timed(lambda i: run_shard(False), args.steps, args.warmup, world, dist, dev, local)
    frames_total = sum(lengths)
    value = frames_total * args.steps / (total_ms / 1e3)
    out_host = torch.empty(sum(lengths[i] for i in mine), 7).pin_memory()

    # end to end: every video's crops / log-mel / BERT rows come from pinned host memory; the copy of
    # video i+1 overlaps the compute of video i (pipeline.HostPrefetcher), logits go back to the host
    from feature_vs_text_compound_emotion_b200.pipeline import HostPrefetcher
    pf = HostPrefetcher(dev)
    offs, o = {}, 0
    for k in mine:
        offs[k] = o
        o += lengths[k]

    def e2e(i):
        local_out = {}

        def one(b):
            return windowing.infer_video(m, b["raw"], {"logmel": b["logmel"], "bert": b["bert"]})

        def sink(j, out):
            k = mine[j]
            local_out[k] = out
            out_host[offs[k]:offs[k] + lengths[k]].copy_(out, non_blocking=True)

        pf.run(({"raw": raw_host[:lengths[k]], "logmel": lm_host[:lengths[k]], "bert": bert_host[:lengths[k]]} for k in mine), one, sink)
        sharding.gather_predictions(local_out, lengths, 7, dev)

    e2e_ms, _ = _timed(e2e, max(1, args.steps // 2), 1, world, dist, dev, local)
    e2e_value = frames_total * max(1, args.steps // 2) / (e2e_ms / 1e3)
    if rank == 0:
        n_mine = sum(lengths[i] for i in mine)
        windows = sum(len(windowing.window_starts(t)) for t in lengths)
        pred = windowing.video_level_prediction(votes[mine[0]])
        burst, sustained, hbm, src = _peaks()
        flops = (IR50_GFLOP_PER_FRAME + 1.7278 + HEAD_MFLOP_PER_FRAME * 1e-3) * 1e9
        line = {
            "metric": "frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {n_videos} videos, T_v~U[150,3000] (seed 7), {frames_total} unique frames, "
                                   f"{windows} windows of 300/hop 200; stored uint8 256x256 crops -> device eval transform -> IR-50; "
                                   "log-mel 96x64 -> VGGish; BERT 768-d; TCN; fusion; stitch; video vote",
                       "sharding": "videos, longest first", "l2": "each video streams >100 MB of crops and >1 GB of activations",
                       "collective": "all_gather of per-frame logits" if world > 1 else "none",
                       "example_vote": pred},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": n_mine * (256 * 256 * 3 + 96 * 64 * 4 + 768 * 4),
                    "d2h_bytes_per_step": n_mine * 28},
            "gpu_launches": None,
            "roofline": {"bound": "tensor", "achieved": value / world * flops / 1e12, "peak": sustained, "unit": "TFLOP/s",
                         "frac": value / world * flops / 1e12 / sustained, "traffic": None,
                         "kernel": "whole pipeline, algorithmic 7.792 GFLOP per unique frame (IR-50 6.0545 + VGGish 1.7278 + head 0.010)",
                         "peak_source": f"{src} sustained"},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_train(args, dev, world, rank, local, dist):
    """configs[3]: fusion-head training step on feature windows (frozen backbones => pre-extracted
    512/128/768-d features), AdamW(lr 1e-4, wd 1e-4), gradients summed with ONE NCCL all-reduce of
    the flat 20 MB bucket; weak scaling (B windows per rank)."""
    from feature_vs_text_compound_emotion_b200 import synthetic
    from feature_vs_text_compound_emotion_b200.models.model import LFAN
    from feature_vs_text_compound_emotion_b200.training import HeadTrainer
    torch.set_grad_enabled(True)
    mods = ["cnn_res50", "vggish", "bert"]
    B = args.train_batch
    m = LFAN(backbone_settings=BS, output_dim=7, task="CLASSIFICATION", modality=mods, kernel_size=5, example_length=LENGTH,
             tcn_channel=synthetic.TCN_CHANNELS, modal_dim=32, num_heads=2, root_dir="", device=dev)
    m.init()
    m.load_state_dict(synthetic.lfan_state_dict(0, mods), strict=True)
    m = m.to(dev).train()
    tr = HeadTrainer(m, B, LENGTH, optimizer={"name": "adamw", "lr": 1e-4, "weight_decay": 1e-4}, seed=rank)
    host = [synthetic.feature_windows(B, LENGTH, seed=200 + i + 100 * rank, modalities=mods) for i in range(2)]
    labs = [torch.randint(0, 7, (B, LENGTH, 1), generator=torch.Generator().manual_seed(300 + i + 100 * rank)) for i in range(2)]
    devb = [{k: v.to(dev) for k, v in h.items()} for h in host]
    devl = [l.to(dev) for l in labs]
    losses = []
    total_ms, clocks = _timed(lambda i: losses.append(tr.step(devb[i % 2], devl[i % 2])), args.steps, args.warmup, world, dist, dev, local)
    frames = B * LENGTH
    value = world * frames * args.steps / (total_ms / 1e3)
    pinned = [{k: v.pin_memory() for k, v in h.items()} for h in host]
    pl = [l.pin_memory() for l in labs]
    loss_host = torch.empty(1).pin_memory()

    # features + labels of step i+1 are staged on a side stream while step i computes
    from feature_vs_text_compound_emotion_b200.pipeline import HostPrefetcher
    pf = HostPrefetcher(dev)

    def e2e_run(n):
        batches = ({**pinned[i % 2], "labels": pl[i % 2]} for i in range(n))
        pf.run(batches, lambda b: tr.step({k: v for k, v in b.items() if k != "labels"}, b["labels"]),
               lambda i, loss: loss_host.copy_(loss, non_blocking=True))

    e2e_run(2)
    e2e_ms, _ = _timed(lambda i: e2e_run(args.steps) if i == 0 else None, 1, 0, world, dist, dev, local)
    if rank == 0:
        burst, sustained, hbm, src = _peaks()
        h2d = sum(v.numel() * v.element_size() for v in pinned[0].values()) + pl[0].numel() * 8
        tf = 3 * HEAD_MFLOP_PER_FRAME * 1e6 * frames / (total_ms / args.steps * 1e-3) / 1e12
        line = {
            "metric": "frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"train: fusion-head step (fwd+CE+bwd+AdamW), {B} windows x {LENGTH} frames per GPU, "
                                   "features 512/128/768-d, dropout 0.1, BatchNorm1d batch stats",
                       "l2": "two rotating batches; 20 MB weights + ~120 MB of saved activations per step",
                       "collective": "one all_reduce(SUM) of the flat 5,002,503-float gradient bucket" if world > 1 else "none",
                       "loss_first_last": [float(losses[0]), float(losses[-1])]},
            "clocks": clocks,
            "e2e": {"value": world * frames * args.steps / (e2e_ms / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": None,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": 72.0, "unit": "TFLOP/s", "frac": tf / 72.0, "traffic": None,
                         "kernel": "row_gemm/wgrad (fp32 CUDA cores; peak = 148 SMs x 128 FMA x 2 x 1.9 GHz), algorithmic 3 x 9.988 MFLOP per frame",
                         "peak_source": "nominal fp32 FMA rate"},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="infer", choices=["infer", "full", "sweep", "train"],
                    help="infer: BASELINE configs[1] shape through the whole path (default, the headline line); "
                         "full: configs[2] -- 56 variable-length videos from stored uint8 crops + log-mel + BERT; "
                         "sweep: configs[4] -- 1000 videos sharded over the ranks (strong scaling); "
                         "train: configs[3] -- fusion-head training step, B=16 windows per rank, NCCL grad all-reduce")
    ap.add_argument("--videos", type=int, default=0, help="override the number of videos of full/sweep")
    ap.add_argument("--train-batch", type=int, default=16)
    ap.add_argument("--frames-per-pass", type=int, default=int(os.environ.get("CER_FRAMES_PER_PASS", "0")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    torch.set_grad_enabled(False)

    from feature_vs_text_compound_emotion_b200 import modules
    if args.frames_per_pass > 0:
        modules.Backbone.frames_per_pass = args.frames_per_pass
    if args.workload in ("full", "sweep"):
        return run_videos(args, dev, world, rank, local, dist)
    if args.workload == "train":
        return run_train(args, dev, world, rank, local, dist)
    model = build_model(dev)
    frames = WINDOWS * LENGTH
    ROT = 4   # rotating input sets: 4 x 54.7 MB > L2, and each step streams > 1 GB of activations
    host = [host_batch(100 + 10 * i + 1000 * rank) for i in range(ROT)]
    devb = [{k: v.to(dev) for k, v in h.items()} for h in host]
    gather = torch.empty(world * frames, 7, device=dev) if world > 1 else None

    def step(batch):
        out = model(dict(batch))
        if world > 1:
            dist.all_gather_into_tensor(gather, out.view(frames, 7))
        return out

    for i in range(args.warmup):
        step(devb[i % ROT])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(devb[i % ROT])
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = ms.item()
    ms_per_step = total_ms / args.steps
    value = world * frames * args.steps / (total_ms / 1e3)

    # ---- IR-50 alone inside the step (device events), for the tensor-pipe fraction of the whole backbone
    vid = devb[0]["video"].view(frames, 3, 40, 40)
    vb = model.spatial["visual"]
    vb(vid)
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(3):
        vb(vid)
    a1.record()
    torch.cuda.synchronize()
    ir50_ms = a0.elapsed_time(a1) / 3

    # ---- end to end through the public API: pinned host -> device, forward, logits -> host
    pinned = [{k: v.pin_memory() for k, v in h.items()} for h in host]
    out_host = torch.empty(WINDOWS, LENGTH, 7).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in pinned[0].values())
    d2h = out_host.numel() * out_host.element_size()

    # the H2D copy of step i+1 runs on a side stream while step i computes (pipeline.HostPrefetcher);
    # every step's copies and its D2H read are inside the timed region
    from feature_vs_text_compound_emotion_b200.pipeline import HostPrefetcher
    pf = HostPrefetcher(dev)

    def e2e_run(n):
        pf.run((pinned[i % ROT] for i in range(n)), step, lambda i, out: out_host.copy_(out, non_blocking=True))

    e2e_run(2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record()
    e2e_run(args.steps)
    b1.record()
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([b0.elapsed_time(b1)], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * frames * args.steps / (e2e_ms.item() / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    burst, sustained, hbm, src = _peaks()
    fpp = modules.Backbone.frames_per_pass
    dom_frames = min(fpp, frames)
    dom_ms, dom_flops = time_dominant_conv(dev, dom_frames)
    achieved = dom_flops / (dom_ms * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "dominant_conv_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    eng = vb.backbone.engine()
    tcn_engines, _ = model._head_engines()
    launches = (eng.launches(frames) + sum(t.launches for t in tcn_engines.values()) + 1) * args.steps
    line = {
        "metric": "frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"LFAN inference (IR-50+TCN+fusion), {WINDOWS} windows x {LENGTH} frames of 40x40 crops per GPU",
                   "frames_per_step_per_gpu": frames, "frames_per_pass": fpp, "head_dtype": "f32",
                   "l2": f"inputs rotate over {ROT} buffers (219 MB > L2); each step streams >1 GB of activations",
                   "collective": "all_gather of per-frame logits" if world > 1 else "none"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst,
                     "traffic": traffic, "kernel": "conv_igemm2_kernel<256,6> (tcgen05 cta_group::2) 256->256 3x3 @10x10",
                     "frames_per_launch": dom_frames, "ms_per_launch": dom_ms, "peak_source": f"{src} burst"},
        "ir50": {"ms": ir50_ms, "tflops": IR50_GFLOP_PER_FRAME * frames / ir50_ms,
                 "frac_of_sustained_peak": IR50_GFLOP_PER_FRAME * frames / ir50_ms / sustained,
                 "share_of_step": ir50_ms / ms_per_step},
    }
    if world == 1 and not args.no_cpu_baseline:
        from feature_vs_text_compound_emotion_b200 import synthetic
        threads = os.cpu_count() or 1
        sd = synthetic.lfan_state_dict(0, MODS)
        t_cal = cpu_reference_step(sd, 16, threads)
        # ~10-30 s of CPU work: as much of one step (2400 frames) as fits, in chunks of 300 frames
        n = int(max(96, min(frames, 300 * max(1, int(20.0 / max(t_cal * 300 / 16, 1e-3))))))
        dt = sum(cpu_reference_step(sd, min(300, n - f0), threads) for f0 in range(0, n, 300))
        line["cpu_baseline"] = {"value": n / dt, "unit": "frames/s", "cores": threads, "kind": "port",
                                "sample": f"{n} frames through IR-50 (chunks of 300) + a {LENGTH}-frame head window per chunk, "
                                          "oracle port fp32, torch CPU on all host threads"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
