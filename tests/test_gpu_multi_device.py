"""One process, two GPUs: the library must work on cuda:1 after cuda:0 (kernel attributes such as the
dynamic shared-memory limit are per device context).  Skipped on single-GPU boxes."""
import warnings

import pytest
import torch

from feature_vs_text_compound_emotion_b200 import synthetic
from oracle import lfan_oracle as O

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")
torch.set_grad_enabled(False)
MODS = ["video", "vggish", "bert"]


def test_second_device_in_the_same_process():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    from feature_vs_text_compound_emotion_b200.models.model import LFAN
    sd = synthetic.lfan_state_dict(0, MODS)
    x = synthetic.frames(300, seed=3).view(1, 300, 3, 40, 40)
    f = synthetic.feature_windows(1, 300, seed=4, modalities=["vggish", "bert"])
    ref = O.lfan_forward(sd, {"video": x, "vggish": f["vggish"], "bert": f["bert"]}, MODS)
    for d in range(2):
        dev = torch.device(f"cuda:{d}")
        m = LFAN(backbone_settings={}, output_dim=7, task="CLASSIFICATION", modality=MODS, kernel_size=5, example_length=300,
                 tcn_channel=synthetic.TCN_CHANNELS, modal_dim=32, num_heads=2, root_dir="", device=dev)
        m.init(visual_state_dict=synthetic.visual_backbone_state_dict(0))
        m.load_state_dict(sd, strict=True)
        m = m.to(dev).eval()
        out = m({"video": x.to(dev), "vggish": f["vggish"].to(dev), "bert": f["bert"].to(dev)}).cpu()
        assert (out - ref).abs().max().item() <= 2e-2, d
