"""VGGish plan alone: patches/s and TFLOP/s at N patches (default 2400), CUDA events.
usage: python tools/vggish_bench.py [patches] [iters]"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from feature_vs_text_compound_emotion_b200 import packing, synthetic
from feature_vs_text_compound_emotion_b200.engine import VggishEngine

torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2400
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
eng = VggishEngine(packing.pack_vggish(synthetic.vggish_state_dict(0)), dev, patches_per_pass=min(n, 2400))
x = synthetic.logmel_patches(n, seed=1).to(dev)
for _ in range(2):
    y = eng.forward(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    y = eng.forward(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"vggish {n} patches: {ms:.3f} ms  {n / ms * 1e3:.0f} patches/s  {1.7278 * n / ms:.1f} TFLOP/s  launches {eng.launches(n)}  mean|y| {float(y.abs().mean()):.3f}")
