"""Device eval transform alone: frames/s and achieved GB/s at N stored 256x256x3 uint8 crops.
usage: python tools/preproc_bench.py [frames] [iters]"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from feature_vs_text_compound_emotion_b200.engine import PreprocEngine

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2400
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
x = torch.randint(0, 256, (n, 256, 256, 3), dtype=torch.uint8, device=dev)
eng = PreprocEngine(256, 256, dev)
out = torch.empty(n, 3, 40, 40, device=dev)
for _ in range(2):
    eng.forward(x, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    eng.forward(x, out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
foot = 219 * 219 * 3 + 19200          # bytes of the crop window's footprint read + fp32 output written, per frame
print(f"preproc {n} frames: {ms:.3f} ms  {n / ms * 1e3:.0f} frames/s  footprint {foot * n / ms / 1e6:.0f} GB/s  "
      f"(whole frames {196608 * n / ms / 1e6:.0f} GB/s)")
