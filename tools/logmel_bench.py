"""Log-mel front end alone: 80 s of 16 kHz audio -> 2400 per-frame examples (30 fps), CUDA events."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
from feature_vs_text_compound_emotion_b200 import synthetic
from feature_vs_text_compound_emotion_b200.engine import LogMelEngine

dev = torch.device("cuda:0")
wave = synthetic.waveform(81.0, seed=1).to(dev)
eng = LogMelEngine(dev)
for _ in range(2):
    ex = eng.examples(wave, 0.96, 1 / 30)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ex = eng.examples(wave, 0.96, 1 / 30)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"logmel: {wave.numel() / 16000:.0f} s audio -> {tuple(ex.shape)} in {ms:.3f} ms  ({ex.shape[0] / ms * 1e3:.0f} examples/s)")
