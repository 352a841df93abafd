"""Head training step alone: steps/s and frames/s at batch B (default 16) x 300, CUDA events.
usage: python tools/train_bench.py [B] [iters]"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from feature_vs_text_compound_emotion_b200 import synthetic
from feature_vs_text_compound_emotion_b200.models.model import LFAN
from feature_vs_text_compound_emotion_b200.training import HeadTrainer

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
mods = ["cnn_res50", "vggish", "bert"]
m = LFAN(backbone_settings={}, output_dim=7, task="CLASSIFICATION", modality=mods, kernel_size=5, example_length=300,
         tcn_channel=synthetic.TCN_CHANNELS, modal_dim=32, num_heads=2, root_dir="", device=dev)
m.init()
m.load_state_dict(synthetic.lfan_state_dict(0, mods), strict=True)
m = m.to(dev).train()
tr = HeadTrainer(m, B, 300, optimizer={"name": "adamw", "lr": 1e-4, "weight_decay": 1e-4})
X = {k: v.to(dev) for k, v in synthetic.feature_windows(B, 300, seed=1, modalities=mods).items()}
y = torch.randint(0, 7, (B, 300, 1)).to(dev)
for _ in range(3):
    loss = tr.step(X, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    loss = tr.step(X, y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"train step B={B}: {ms:.3f} ms  {B * 300 / ms * 1e3:.0f} frames/s  {29.96e-3 * B * 300 / ms:.2f} TFLOP/s  loss {loss.item():.4f}")
