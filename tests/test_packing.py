"""Host logic: the packed-weight algebra reproduces the oracle (CPU, fp32 / emulated bf16)."""
import torch

from feature_vs_text_compound_emotion_b200 import packing, synthetic
from oracle import lfan_oracle as O
from tests import emulate

torch.set_grad_enabled(False)


def test_ir50_packing_exact_in_fp32():
    sd = synthetic.visual_backbone_state_dict(0)
    pk = packing.pack_ir50(sd, "backbone.")
    x = synthetic.frames(2, seed=5)
    ref = O.ir50_forward(sd, x, "backbone.")
    emb, _ = emulate.ir50_packed(pk, x, round_act=False)
    # weights are bf16-rounded in `pk`, activations are not: error is small but not fp32-exact
    cos = torch.nn.functional.cosine_similarity(emb, ref, dim=1)
    assert cos.min().item() > 0.9999


def test_ir50_border_bias_and_projection_exact():
    """Same check with fp32 weights (no bf16 anywhere) => agreement to fp32 round-off, which
    proves the 9-class border table, the fused projection shortcut and the FC permutation."""
    sd = synthetic.visual_backbone_state_dict(1)
    pk = packing.pack_ir50(sd, "backbone.")
    pk32 = packing.pack_ir50(sd, "backbone.", operand_dtype=torch.float32)
    x = synthetic.frames(2, seed=6)
    ref = O.ir50_forward(sd, x, "backbone.")
    emb, _ = emulate.ir50_packed(pk32, x, round_act=False)
    assert (emb - ref).abs().max().item() < 2e-5
    assert pk["units"][3]["has_proj"] == 1 and pk["units"][3]["w2"].shape == (128, 9 * 128 + 64)


def test_ir50_bf16_budget():
    """bf16 operands + bf16 residual stream: predicted cosine vs fp32 oracle (budget 0.999)."""
    sd = synthetic.visual_backbone_state_dict(0)
    pk = packing.pack_ir50(sd, "backbone.")
    x = synthetic.frames(4, seed=1234)
    ref = O.ir50_forward(sd, x, "backbone.")
    emb, _ = emulate.ir50_packed(pk, x, round_act=True)
    cos = torch.nn.functional.cosine_similarity(emb, ref, dim=1)
    assert cos.min().item() > 0.9995, cos


def test_tcn_and_fusion_packing():
    mods = ["cnn_res50", "vggish", "bert"]
    sd = synthetic.lfan_state_dict(0, mods)
    X = synthetic.feature_windows(2, 300, seed=3, modalities=mods)
    ref = O.lfan_forward(sd, X, mods)
    enc = []
    for m in mods:
        blocks = packing.pack_tcn(sd, f"temporal.{m}.", f"bn.{m}")
        enc.append(emulate.tcn_packed(blocks, X[m].squeeze(1)))
        want = O._bn_eval(sd, f"bn.{m}", O.tcn_forward(sd, f"temporal.{m}.", X[m].squeeze(1).transpose(1, 2))).transpose(1, 2)
        assert (enc[-1] - want).abs().max().item() < 1e-4
    fw = packing.pack_fusion(sd, mods, 32, 2)
    logits, _ = emulate.fusion_packed(fw, [e.reshape(600, -1) for e in enc])
    assert (logits.view(2, 300, 7) - ref).abs().max().item() < 1e-4


def test_padded_raster_addressing_equals_zero_padded_conv():
    """The layout algebra of csrc/conv_raster.cuh on the CPU: store a batch of maps as a padded raster
    [frames][H+1][W+1][C] (one zero column after every row, one zero row after every frame); then for EVERY
    position q the input of tap (r, s) is the element at q + (r-1)*(W+1) + (s-1) (zero before / after the
    tensor), so a 3x3 / pad 1 convolution is nine shifted views of ONE flat tensor -- frame borders included."""
    g = torch.Generator().manual_seed(5)
    n, H, W, C, Co = 3, 5, 4, 2, 3
    x = torch.randn(n, C, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(Co, C, 3, 3, generator=g, dtype=torch.float64)
    want = torch.nn.functional.conv2d(x, w, padding=1)                       # [n, Co, H, W]
    wp, P = W + 1, (H + 1) * (W + 1)
    ras = torch.zeros(n, H + 1, wp, C, dtype=torch.float64)
    ras[:, :H, :W] = x.permute(0, 2, 3, 1)
    flat = ras.reshape(n * P, C)
    lead = wp + 1                                                            # the most negative offset is -(wp + 1)
    padded = torch.cat([torch.zeros(lead, C, dtype=torch.float64), flat, torch.zeros(lead, C, dtype=torch.float64)])
    out = torch.zeros(n * P, Co, dtype=torch.float64)
    for r in range(3):
        for s in range(3):
            off = (r - 1) * wp + (s - 1)
            view = padded[lead + off: lead + off + n * P]                    # the whole tensor shifted by one tap
            out += view @ w[:, :, r, s].T
    got = out.reshape(n, H + 1, wp, Co)[:, :H, :W].permute(0, 3, 1, 2)
    assert torch.allclose(got, want, atol=1e-12)
    # the kernel's position -> pixel map and the number of pad positions per frame (raster_zero_pads_kernel)
    q = torch.arange(n * P)
    rem = q % P
    real = (rem // wp < H) & (rem % wp < W)
    assert int(real.sum()) == n * H * W and int((~real).sum()) == n * (H + wp)
