"""Does IR-50 stage 1 get faster when its activations fit in L2?  Times ops [first, last] of the plan
(cer_ir50_run_ops) at several pass sizes and prints microseconds per frame.  usage: l2_chunk_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from feature_vs_text_compound_emotion_b200 import packing, synthetic  # noqa: E402
from feature_vs_text_compound_emotion_b200.engine import Ir50Engine  # noqa: E402

dev = torch.device("cuda:0")
pk = packing.pack_ir50(synthetic.visual_backbone_state_dict(0), "backbone.")
eng = Ir50Engine(pk, dev, frames_per_pass=2400)
x = torch.randn(2400, 3, 40, 40, device=dev)
groups = {"stem": (0, 0), "stage1 (6 convs)": (1, 6), "unit3 (widen + s2)": (7, 8), "stage2 (6 convs)": (9, 14),
          "stem..stage2": (0, 14), "stage3 (26 convs)": (17, 42)}
for frames in (100, 150, 200, 300, 600, 1200, 2400):
    eng.forward(x[:frames])
    row = []
    for name, (a, b) in groups.items():
        for _ in range(3):
            eng.run_ops(x, frames, a, b)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(5, 4800 // frames)
        e0.record()
        for _ in range(reps):
            eng.run_ops(x, frames, a, b)
        e1.record()
        torch.cuda.synchronize()
        row.append(f"{name}: {e0.elapsed_time(e1) / reps / frames * 1e3:.3f}")
    print(f"frames {frames:5d} | us/frame | " + " | ".join(row) + f" | variant(op0)={eng.op_variant(0, frames)} op8={eng.op_variant(8, frames)}")
