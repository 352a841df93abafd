"""Per-layer-class timing of the tcgen05 implicit-GEMM conv through cer_conv_forward.
usage: python tools/conv_sweep.py [frames ...]"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from feature_vs_text_compound_emotion_b200.engine import conv_forward

dev = torch.device("cuda:0")
CLASSES = [  # name, H, cin, cout, stride, classes, prelu, residual
    ("s1 64->64 @40", 40, 64, 64, 1, 9, True, False),
    ("s1 64->64 @40 +res", 40, 64, 64, 1, 1, False, True),
    ("w 64->128 @40", 40, 64, 128, 1, 9, True, False),
    ("s2 128->128 @20", 20, 128, 128, 1, 9, True, False),
    ("w 128->256 @20", 20, 128, 256, 1, 9, True, False),
    ("s3 256->256 @10", 10, 256, 256, 1, 9, True, False),
    ("s3 256->256 @10 +res", 10, 256, 256, 1, 1, False, True),
    ("s4 512->512 @5", 5, 512, 512, 1, 9, True, False),
]
frames_list = [int(a) for a in sys.argv[1:]] or [74, 296, 2400]
g = torch.Generator().manual_seed(0)
for name, H, cin, cout, stride, classes, prelu, residual in CLASSES:
    for frames in frames_list:
        x = torch.randn(frames + 8, H, H, cin, generator=g).to(torch.bfloat16).to(dev)
        w = (torch.randn(cout, 9 * cin, generator=g) * (9 * cin) ** -0.5).to(torch.bfloat16).to(dev)
        bias = torch.randn(classes, cout, generator=g).to(dev)
        alpha = torch.full((cout,), 0.25).to(dev) if prelu else None
        res = torch.randn(frames, H, H, cout, generator=g).to(torch.bfloat16).to(dev) if residual else None
        for _ in range(3):
            conv_forward(x, w, bias, 3, stride, 1, alpha=alpha, res=res, n_frames=frames)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            conv_forward(x, w, bias, 3, stride, 1, alpha=alpha, res=res, n_frames=frames)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = 2.0 * frames * H * H * cout * 9 * cin
        tiles = (frames * H * H + 127) // 128 * max(1, cout // 256)
        print(f"{name:24s} frames {frames:5d} tiles {tiles:6d}  {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s", flush=True)
