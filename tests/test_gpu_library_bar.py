"""The "library bar" of SURVEY.md section 8d: the reference arithmetic (oracle restatement, i.e. the
same F.conv2d / F.linear calls the reference's modules make) run by eager PyTorch + cuDNN/cuBLAS on
the SAME B200, in fp32 (TF32 off), fp32 with TF32 on, and bf16 autocast, next to this repo's
kernels.  Not a parity test: it records numbers in gpurun_out/library_bar.json and only asserts
that the hand-written path is not slower than the library path it replaces."""
import json
import os
import warnings

import pytest
import torch

from feature_vs_text_compound_emotion_b200 import synthetic
from oracle import lfan_oracle as O

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")
torch.set_grad_enabled(False)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _time(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def test_ir50_library_bar():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    dev = torch.device("cuda:0")
    from feature_vs_text_compound_emotion_b200.models.backbone import VisualBackbone
    n = 2400
    sd = synthetic.visual_backbone_state_dict(0)
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    x = synthetic.frames(n, seed=9).to(dev)
    res = {"frames": n}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        res["eager_fp32_ms"] = _time(lambda: O.ir50_forward(sd_dev, x, "backbone."))
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        res["eager_tf32_ms"] = _time(lambda: O.ir50_forward(sd_dev, x, "backbone."))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            res["eager_bf16_autocast_ms"] = _time(lambda: O.ir50_forward(sd_dev, x, "backbone."))
        with torch.autocast("cuda", dtype=torch.float16):          # the reference's --amp True (trainer.py:367,478)
            res["eager_fp16_autocast_ms"] = _time(lambda: O.ir50_forward(sd_dev, x, "backbone."))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    vb = VisualBackbone(use_pretrained=False)
    vb.load_state_dict(sd, strict=True)
    vb = vb.to(dev).eval()
    res["b200_kernels_ms"] = _time(lambda: vb(x), reps=10)
    for k in list(res):
        if k.endswith("_ms"):
            res[k.replace("_ms", "_frames_per_s")] = n / res[k] * 1e3
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "library_bar.json"), "w"), indent=1)
    print(json.dumps(res))
    assert res["b200_kernels_ms"] < res["eager_bf16_autocast_ms"]
