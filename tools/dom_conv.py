"""The dominant conv class (256->256 3x3 @10x10, PReLU epilogue, 2400 frames) alone through
cer_conv_forward -- the command `ncu --set full` wraps for roofline.traffic.
usage: python tools/dom_conv.py [launches] [H cin cout]"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from feature_vs_text_compound_emotion_b200.engine import conv_forward

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H, cin, cout = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (10, 256, 256)
frames = 2400
g = torch.Generator().manual_seed(0)
x = torch.randn(frames + 8, H, H, cin, generator=g).to(torch.bfloat16).to(dev)
w = (torch.randn(cout, 9 * cin, generator=g) * (9 * cin) ** -0.5).to(torch.bfloat16).to(dev)
bias = torch.randn(9, cout, generator=g).to(dev)
alpha = torch.full((cout,), 0.25).to(dev)
for _ in range(n):
    y = conv_forward(x, w, bias, 3, 1, 1, alpha=alpha, n_frames=frames)
torch.cuda.synchronize()
print("ok", tuple(y.shape), float(y.float().abs().mean()))
