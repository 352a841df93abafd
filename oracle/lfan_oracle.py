"""CPU oracle for the LFAN hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch fp32, *functional* restatement of the reference
arithmetic for the path named in BASELINE.json (IR-50 -> TCN -> cross-modal
attention -> classifier).  It works directly on a reference-layout
``state_dict`` (dict[str, Tensor]); it holds no nn.Module and no parameters.

Who may use it: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product
package (``feature_vs_text_compound_emotion_b200``) never imports it.

Parity pin: the reference ships no golden vectors, so the
oracle is pinned against the reference modules themselves, imported from
/root/reference in the build container by ``oracle/gen_golden.py``; the
resulting fixtures live in ``tests/golden/`` and ``tests/test_oracle.py``
re-checks the oracle against them on every run (no /root/reference needed).

Every function cites the reference lines it restates (paths relative to
/root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

BN_EPS = 1e-5          # torch.nn.BatchNorm{1,2}d default, used everywhere in the reference
LN_EPS = 1e-5          # torch.nn.LayerNorm default (models/transformer.py:189)
LEAKY_SLOPE = 0.01     # torch.nn.LeakyReLU default (models/temporal_convolutional_model.py:27,33,39)

# get_blocks(50): models/arcface_model.py:95-102 -> (in_channel, depth, stride) per unit
IR50_UNITS = ([(64, 64, 1)] * 3
              + [(64, 128, 2)] + [(128, 128, 1)] * 3
              + [(128, 256, 2)] + [(256, 256, 1)] * 13
              + [(256, 512, 2)] + [(512, 512, 1)] * 2)


def _bn_eval(sd: SD, p: str, x: Tensor) -> Tensor:
    """Eval-mode BatchNorm as an affine map over dim 1 (nn.BatchNorm1d/2d, running stats)."""
    scale = sd[p + ".weight"] / torch.sqrt(sd[p + ".running_var"] + BN_EPS)
    shift = sd[p + ".bias"] - sd[p + ".running_mean"] * scale
    shape = [1, -1] + [1] * (x.dim() - 2)
    return x * scale.view(shape) + shift.view(shape)


def _prelu(x: Tensor, alpha: Tensor) -> Tensor:
    """Per-channel PReLU (models/arcface_model.py:54, :132)."""
    a = alpha.view(1, -1, 1, 1)
    return torch.where(x >= 0, x, a * x)


# --------------------------------------------------------------------------------------
# IR-50 (models/arcface_model.py:120-151, models/backbone.py:69-130)
# --------------------------------------------------------------------------------------
def ir_unit(sd: SD, p: str, x: Tensor, cin: int, depth: int, stride: int) -> Tensor:
    """bottleneck_IR.forward, models/arcface_model.py:44-60."""
    if cin == depth:
        # MaxPool2d(1, stride) == strided subsampling (:48)
        shortcut = x[:, :, ::stride, ::stride]
    else:
        shortcut = F.conv2d(x, sd[p + ".shortcut_layer.0.weight"], None, stride)          # :50
        shortcut = _bn_eval(sd, p + ".shortcut_layer.1", shortcut)                          # :51
    r = _bn_eval(sd, p + ".res_layer.0", x)                                                 # :53
    r = F.conv2d(r, sd[p + ".res_layer.1.weight"], None, 1, 1)                              # :54
    r = _prelu(r, sd[p + ".res_layer.2.weight"])                                            # :54
    r = F.conv2d(r, sd[p + ".res_layer.3.weight"], None, stride, 1)                         # :55
    r = _bn_eval(sd, p + ".res_layer.4", r)                                                 # :55
    return r + shortcut                                                                     # :60


def ir50_forward(sd: SD, x: Tensor, prefix: str = "", units=IR50_UNITS) -> Tensor:
    """Backbone.forward (arcface_model.py:147-151) with the 5x5 head of VisualBackbone
    (backbone.py:99-103).  ``x`` is [N,3,40,40] fp32; returns unit-norm [N,512].
    ``prefix`` selects the sub-dict, e.g. 'backbone.' for a VisualBackbone state_dict or
    'spatial.visual.backbone.' for an LFAN one."""
    p = prefix
    h = F.conv2d(x, sd[p + "input_layer.0.weight"], None, 1, 1)                             # :130
    h = _bn_eval(sd, p + "input_layer.1", h)                                                # :131
    h = _prelu(h, sd[p + "input_layer.2.weight"])                                           # :132
    for i, (cin, depth, stride) in enumerate(units):
        h = ir_unit(sd, f"{p}body.{i}", h, cin, depth, stride)
    h = _bn_eval(sd, p + "output_layer.0", h)            # BN2d; Dropout is identity in eval
    h = h.reshape(h.shape[0], -1)                         # Flatten in NCHW order (arcface_model.py:12-14)
    h = F.linear(h, sd[p + "output_layer.3.weight"], sd[p + "output_layer.3.bias"])
    h = _bn_eval(sd, p + "output_layer.4", h)             # BN1d
    return h / torch.norm(h, 2, 1, True)                  # l2_norm, arcface_model.py:17-20


# --------------------------------------------------------------------------------------
# VGGish audio backbone (models/backbone.py:16-66, :133-145)
# --------------------------------------------------------------------------------------
VGGISH_CFG = (64, "M", 128, "M", 256, 256, "M", 512, 512, "M")       # make_layers(), backbone.py:43-53
VGGISH_CONV_IDX = (0, 3, 6, 8, 11, 13)                                # positions of the Conv2d's in `features`


def vggish_forward(sd: SD, x: Tensor, prefix: str = "") -> Tensor:
    """VGGish.forward -> VGG.forward (backbone.py:29-40, :63-66).  x: [N, 96, 64] log-mel patches
    (the reference wraps them as [N,1,96,64]); returns [N,128].  Conv3x3(pad 1)+ReLU stacks with
    2x2 max-pools, the feature map is flattened in (h, w, c) order (:34-37) and goes through
    Linear-ReLU-Linear-ReLU-Linear (no final ReLU, :20-27)."""
    h = x[:, None, :, :].float()                                                   # :64
    li = 0
    for v in VGGISH_CFG:
        if v == "M":
            h = F.max_pool2d(h, 2, 2)
        else:
            i = VGGISH_CONV_IDX[li]
            h = F.relu(F.conv2d(h, sd[f"{prefix}features.{i}.weight"], sd[f"{prefix}features.{i}.bias"], 1, 1))
            li += 1
    h = h.transpose(1, 3).transpose(1, 2).contiguous()                             # NCHW -> NHWC, :34-36
    h = h.view(h.shape[0], -1)
    h = F.relu(F.linear(h, sd[prefix + "embeddings.0.weight"], sd[prefix + "embeddings.0.bias"]))
    h = F.relu(F.linear(h, sd[prefix + "embeddings.2.weight"], sd[prefix + "embeddings.2.bias"]))
    return F.linear(h, sd[prefix + "embeddings.4.weight"], sd[prefix + "embeddings.4.bias"])


# --------------------------------------------------------------------------------------
# TCN (models/temporal_convolutional_model.py:12-75)
# --------------------------------------------------------------------------------------
def weight_norm_effective(g: Tensor, v: Tensor) -> Tensor:
    """Old-style torch.nn.utils.weight_norm (dim=0): w = g * v / ||v|| with the norm taken
    over every dim but 0 (temporal_convolutional_model.py:24, :30)."""
    norm = v.reshape(v.shape[0], -1).norm(dim=1).view(-1, 1, 1)
    return g * v / norm


def causal_dilated_conv(x: Tensor, w: Tensor, b: Tensor, dilation: int) -> Tensor:
    """Conv1d(padding=(k-1)*d, dilation=d) followed by Chomp1d(padding)
    (temporal_convolutional_model.py:12-18, :24-26): symmetric padding, then drop the
    right-hand ``padding`` samples => causal."""
    pad = (w.shape[-1] - 1) * dilation
    y = F.conv1d(x, w, b, stride=1, padding=pad, dilation=dilation)
    return y[:, :, :-pad] if pad > 0 else y


def temporal_block(sd: SD, p: str, x: Tensor, dilation: int) -> Tensor:
    """TemporalBlock.forward (:51-54); Dropout is identity in eval."""
    w1 = weight_norm_effective(sd[p + ".conv1.weight_g"], sd[p + ".conv1.weight_v"])
    w2 = weight_norm_effective(sd[p + ".conv2.weight_g"], sd[p + ".conv2.weight_v"])
    h = F.leaky_relu(causal_dilated_conv(x, w1, sd[p + ".conv1.bias"], dilation), LEAKY_SLOPE)
    h = F.leaky_relu(causal_dilated_conv(h, w2, sd[p + ".conv2.bias"], dilation), LEAKY_SLOPE)
    if (p + ".downsample.weight") in sd:
        res = F.conv1d(x, sd[p + ".downsample.weight"], sd[p + ".downsample.bias"])
    else:
        res = x
    return F.leaky_relu(h + res, LEAKY_SLOPE)


def tcn_forward(sd: SD, prefix: str, x: Tensor, num_levels: int = 4) -> Tensor:
    """TemporalConvNet.forward (:57-75): level i has dilation 2**i.  x: [B, C_in, T]."""
    for i in range(num_levels):
        x = temporal_block(sd, f"{prefix}network.{i}", x, 2 ** i)
    return x


# --------------------------------------------------------------------------------------
# Cross-modal attention fusion (models/transformer.py:11-19, :102-215)
# --------------------------------------------------------------------------------------
def fusion_forward(sd: SD, prefix: str, feats: Dict[str, Tensor], modalities: Sequence[str],
                   modal_dim: int = 32, num_heads: int = 2) -> Tensor:
    """MultimodalTransformerEncoder.forward -> MultiModalEncoderBlock.forward ->
    MultimodalMultiheadAttention.forward.  feats[m]: [B, T, D_m].  Returns [B, T, modal_dim*M].

    The reference reshapes each qkv projection [B,T,3*modal_dim] as [B,T,H,1,3*hd] and chunks
    the last axis (:142-144), so inside one head the 3*hd outputs are laid out q|k|v.  Tokens
    are the modalities (M of them); softmax(QK^T/sqrt(hd))V + V (:156-157); heads/modalities
    are flattened as (head, modal, dim) (:158-159); o_proj; LayerNorm without residual
    (:192-197, dropout is identity in eval)."""
    hd = modal_dim // num_heads
    a = prefix + "layers.self_attn."
    qs, ks, vs = [], [], []
    for m in modalities:
        qkv = F.linear(feats[m], sd[f"{a}qkv_proj.{m}.weight"], sd[f"{a}qkv_proj.{m}.bias"])
        B, T, _ = qkv.shape
        qkv = qkv.view(B, T, num_heads, 3, hd)
        qs.append(qkv[:, :, :, 0])
        ks.append(qkv[:, :, :, 1])
        vs.append(qkv[:, :, :, 2])
    Q = torch.stack(qs, dim=3)          # [B, T, H, M, hd]
    K = torch.stack(ks, dim=3)
    V = torch.stack(vs, dim=3)
    logits = torch.einsum("bthmd,bthnd->bthmn", Q, K) / math.sqrt(hd)
    att = torch.softmax(logits, dim=-1)
    vals = torch.einsum("bthmn,bthnd->bthmd", att, V) + V
    vals = vals.reshape(B, T, num_heads * len(modalities) * hd)
    o = F.linear(vals, sd[a + "o_proj.weight"], sd[a + "o_proj.bias"])
    n = prefix + "layers.norm1."
    return F.layer_norm(o, (o.shape[-1],), sd[n + "weight"], sd[n + "bias"], LN_EPS)


# --------------------------------------------------------------------------------------
# LFAN (models/model.py:375-526)
# --------------------------------------------------------------------------------------
def head_forward(sd: SD, feats: Dict[str, Tensor], modalities: Sequence[str],
                 modal_dim: int = 32, num_heads: int = 2) -> Tensor:
    """Everything in LFAN.forward after the backbones (model.py:511-526).
    feats[m]: [B, T, D_m] (the reference receives [B,1,T,D_m] and squeezes, :513).
    Returns logits [B, T, n_cls]."""
    enc = {}
    for m in modalities:
        x = feats[m].transpose(1, 2)                                  # :513  [B, D, T]
        x = tcn_forward(sd, f"temporal.{m}.", x)                      # :514
        enc[m] = _bn_eval(sd, f"bn.{m}", x).transpose(1, 2)           # :515  [B, T, C]
    follower = fusion_forward(sd, "fusion.", enc, modalities, modal_dim, num_heads)   # :517
    cat = torch.cat((enc[modalities[0]], follower), dim=-1)           # :519
    return F.linear(cat, sd["regressor.weight"], sd["regressor.bias"])   # :520


def lfan_forward(sd: SD, X: Dict[str, Tensor], modalities: Sequence[str],
                 modal_dim: int = 32, num_heads: int = 2) -> Tensor:
    """LFAN.forward (model.py:487-526), classification task (no tanh, :523).
    X['video']: [B,T,3,40,40]; X['logmel']: [B,64,T,96]; other modalities [B,1,T,D]."""
    feats = {}
    for m in modalities:
        if m == "video":
            B, T = X[m].shape[:2]
            emb = ir50_forward(sd, X[m].reshape(B * T, *X[m].shape[2:]), "spatial.visual.backbone.")
            feats[m] = emb.view(B, T, -1)                             # :490-497
        elif m == "logmel":
            B, hh, T, ww = X[m].shape                                 # :500  [B, 64, T, 96]
            patches = X[m].permute(0, 2, 3, 1).contiguous().view(-1, ww, hh)      # :501-502
            feats[m] = vggish_forward(sd, patches, "spatial.audio.backbone.").view(B, T, -1)   # :504-507
        else:
            feats[m] = X[m].squeeze(1)
    return head_forward(sd, feats, modalities, modal_dim, num_heads)


# --------------------------------------------------------------------------------------
# Long-video windowing (trainer.py:788-913)
# --------------------------------------------------------------------------------------
def windowing(length: int, window_length: int = 300, hop_length: int = 200) -> List[np.ndarray]:
    """Trainer.windowing (trainer.py:894-913) on arange(length): full windows every hop, plus
    a tail window covering the last ``window_length`` frames when the regular grid misses the
    end; a single short window when length < window_length."""
    idx = np.arange(length)
    if length < window_length:
        return [idx]
    n = (length - window_length) // hop_length + 1
    out = [idx[i * hop_length: i * hop_length + window_length] for i in range(n)]
    if out[-1][-1] < length - 1:
        out.append(idx[-window_length:])
    return out


def windowed_inference(forward_fn, X: Dict[str, Tensor], window_length: int = 300,
                       hop_length: int = 200) -> Tensor:
    """Trainer.inference_forward_windows (trainer.py:832-892): run ``forward_fn`` on each
    window, sum into place, divide by the per-frame overlap count."""
    any_key = next(iter(X))
    length = X[any_key].shape[1] if any_key == "video" else X[any_key].shape[2]
    out, cnt = None, torch.zeros(length)
    for wd in windowing(length, window_length, hop_length):
        chunk = {m: (v[:, wd] if m == "video" else v[:, :, wd]) for m, v in X.items()}
        y = forward_fn(chunk)
        if out is None:
            out = torch.zeros(y.shape[0], length, y.shape[2], dtype=y.dtype)
        out[:, wd] += y
        cnt[wd] += 1
    return out / cnt.view(1, -1, 1)


# --------------------------------------------------------------------------------------
# Fusion-head training step (BASELINE config 4; trainer.py:365-391, experiment.py:133)
# --------------------------------------------------------------------------------------
BN_MOMENTUM = 0.1      # torch.nn.BatchNorm1d default (models/model.py:475)
TCN_DROPOUT = 0.1      # models/model.py:471
FUSION_DROPOUT = 0.1   # models/model.py:482


def dropout_keep_mask(shape, p: float, seed: int, stream: int) -> Tensor:
    """The dropout RNG of THIS repo's training kernels (csrc/train.cu `dropout_keep`), restated in
    numpy so the oracle can apply exactly the masks the kernels draw.  It is not torch's Philox
    stream: the reference's masks are not reproducible outside torch, so parity of the training
    step is pinned with p = 0 against the reference and with these masks against the oracle.
    keep(idx) = fmix32(idx * 0x9E3779B1 + seed + stream * 0x85EBCA6B) >= p * 2^32, idx = row-major
    element index of the [rows, channels] activation."""
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        x = np.arange(n, dtype=np.uint32) * np.uint32(0x9E3779B1) + np.uint32((seed + stream * 0x85EBCA6B) & 0xFFFFFFFF)
        x ^= x >> np.uint32(16)
        x *= np.uint32(0x85EBCA6B)
        x ^= x >> np.uint32(13)
        x *= np.uint32(0xC2B2AE35)
        x ^= x >> np.uint32(16)
    thr = np.uint32(min(int(p * 4294967296.0), 0xFFFFFFFF))
    return torch.from_numpy((x >= thr).astype(np.float32)).view(*shape)


def _drop(x: Tensor, p: float, seed, stream: int) -> Tensor:
    """nn.Dropout in training mode on a [B, T, C] (time-major) activation; seed None => p = 0."""
    if seed is None or p <= 0.0:
        return x
    return x * dropout_keep_mask(tuple(x.shape), p, seed, stream) / (1.0 - p)


def dropout_stream(modal_index: int, block: int, site: int) -> int:
    """Stream ids shared with the kernels: TCN dropout1/dropout2 of block i of modality m."""
    return modal_index * 16 + block * 2 + site


FUSION_STREAM = 4096


def tf32_round(x: Tensor) -> Tensor:
    """Round fp32 to TF32 (10 mantissa bits) to nearest, ties away from zero: PTX cvt.rna.tf32.f32."""
    bits = x.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


class _Tf32Conv1d(torch.autograd.Function):
    """conv1d whose three GEMMs -- forward, gradient w.r.t. the input, gradient w.r.t. the weight -- see
    TF32-rounded operands and accumulate in fp32: the arithmetic of the training kernels in
    precision="tf32" (and of cuDNN with torch.backends.cudnn.allow_tf32, the reference's GPU default)."""

    @staticmethod
    def forward(ctx, x, w, b, padding, dilation):
        ctx.save_for_backward(x, w)
        ctx.cfg = (padding, dilation)
        return F.conv1d(tf32_round(x), tf32_round(w), b, stride=1, padding=padding, dilation=dilation)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        padding, dilation = ctx.cfg
        g = tf32_round(gy)
        gx = torch.nn.grad.conv1d_input(x.shape, tf32_round(w), g, stride=1, padding=padding, dilation=dilation)
        gw = torch.nn.grad.conv1d_weight(tf32_round(x), w.shape, g, stride=1, padding=padding, dilation=dilation)
        return gx, gw, gy.sum(dim=(0, 2)), None, None


def _conv1d(x, w, b, padding, dilation, tf32):
    if tf32:
        return _Tf32Conv1d.apply(x, w, b, padding, dilation)
    return F.conv1d(x, w, b, stride=1, padding=padding, dilation=dilation)


def tcn_bn_train(P: SD, buffers: SD, feats: Dict[str, Tensor], modalities: Sequence[str], seed=None,
                 p_tcn: float = TCN_DROPOUT, tf32: bool = False) -> Dict[str, Tensor]:
    """TemporalConvNet + BatchNorm1d of every modality in TRAINING mode (what LFAN, CAN and JMT / MT share:
    model.py:511-515, :672-676, :1155-1159): dropout after both LeakyReLUs of every block, batch statistics,
    running-stat update in ``buffers``.  feats[m]: [B, T, D_m] -> enc[m]: [B, T, C_m]."""
    enc = {}
    for mi, m in enumerate(modalities):
        x = feats[m]
        i = 0
        while f"temporal.{m}.network.{i}.conv1.weight_v" in P:
            p = f"temporal.{m}.network.{i}"
            d = 2 ** i
            w1 = weight_norm_effective(P[p + ".conv1.weight_g"], P[p + ".conv1.weight_v"])
            w2 = weight_norm_effective(P[p + ".conv2.weight_g"], P[p + ".conv2.weight_v"])
            xc = x.transpose(1, 2)
            pad = (w1.shape[-1] - 1) * d

            def cconv(inp, w, bias):           # causal_dilated_conv, optionally with TF32-rounded GEMM operands
                y = _conv1d(inp, w, bias, pad, d, tf32)
                return y[:, :, :-pad] if pad > 0 else y

            h = F.leaky_relu(cconv(xc, w1, P[p + ".conv1.bias"]), LEAKY_SLOPE).transpose(1, 2)
            h = _drop(h, p_tcn, seed, dropout_stream(mi, i, 0))
            h = F.leaky_relu(cconv(h.transpose(1, 2), w2, P[p + ".conv2.bias"]), LEAKY_SLOPE).transpose(1, 2)
            h = _drop(h, p_tcn, seed, dropout_stream(mi, i, 1))
            if (p + ".downsample.weight") in P:
                res = _conv1d(xc, P[p + ".downsample.weight"], P[p + ".downsample.bias"], 0, 1, tf32).transpose(1, 2)
            else:
                res = x
            x = F.leaky_relu(h + res, LEAKY_SLOPE)
            i += 1
        xb = F.batch_norm(x.transpose(1, 2), buffers[f"bn.{m}.running_mean"], buffers[f"bn.{m}.running_var"],
                          P[f"bn.{m}.weight"], P[f"bn.{m}.bias"], True, BN_MOMENTUM, BN_EPS)
        if f"bn.{m}.num_batches_tracked" in buffers:
            buffers[f"bn.{m}.num_batches_tracked"] += 1
        enc[m] = xb.transpose(1, 2)
    return enc


def head_forward_train(P: SD, buffers: SD, feats: Dict[str, Tensor], modalities: Sequence[str],
                       modal_dim: int = 32, num_heads: int = 2, seed=None,
                       p_tcn: float = TCN_DROPOUT, p_fusion: float = FUSION_DROPOUT, tf32: bool = False) -> Tensor:
    """LFAN.forward after the backbones in TRAINING mode (model.py:511-526 under model.train()):
    Dropout active after both LeakyReLUs of every TemporalBlock (temporal_convolutional_model.py:
    28,34) and on the attention output (transformer.py:194); BatchNorm1d uses batch statistics and
    updates ``buffers`` (running_mean/var with the unbiased variance, momentum 0.1).
    P: trainable tensors (reference names), feats[m]: [B, T, D_m].  Activations are kept
    time-major [B, T, C] so that dropout indices match the kernels' layout."""
    enc = tcn_bn_train(P, buffers, feats, modalities, seed, p_tcn, tf32)
    hd = modal_dim // num_heads
    a = "fusion.layers.self_attn."
    qs, ks, vs = [], [], []
    for m in modalities:
        qkv = F.linear(enc[m], P[f"{a}qkv_proj.{m}.weight"], P[f"{a}qkv_proj.{m}.bias"])
        B, T, _ = qkv.shape
        qkv = qkv.view(B, T, num_heads, 3, hd)
        qs.append(qkv[:, :, :, 0]); ks.append(qkv[:, :, :, 1]); vs.append(qkv[:, :, :, 2])
    Q, K, V = torch.stack(qs, 3), torch.stack(ks, 3), torch.stack(vs, 3)
    att = torch.softmax(torch.einsum("bthmd,bthnd->bthmn", Q, K) / math.sqrt(hd), dim=-1)
    vals = (torch.einsum("bthmn,bthnd->bthmd", att, V) + V).reshape(B, T, -1)
    o = F.linear(vals, P[a + "o_proj.weight"], P[a + "o_proj.bias"])
    o = _drop(o, p_fusion, seed, FUSION_STREAM)
    f = F.layer_norm(o, (o.shape[-1],), P["fusion.layers.norm1.weight"], P["fusion.layers.norm1.bias"], LN_EPS)
    cat = torch.cat((enc[modalities[0]], f), dim=-1)
    return F.linear(cat, P["regressor.weight"], P["regressor.bias"])


def trainable_names(sd: SD) -> List[str]:
    """Names of the head's trainable tensors in ``parameters()`` order: everything but spatial.*,
    BatchNorm buffers and the duplicated net.{0,4}.* aliases of conv1/conv2."""
    out = []
    for k in sd:
        if k.startswith("spatial.") or ".net." in k:
            continue
        if k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"):
            continue
        out.append(k)
    return out


def train_step(sd: SD, feats: Dict[str, Tensor], labels: Tensor, modalities: Sequence[str], opt: dict,
               opt_state: dict = None, seed=None, modal_dim: int = 32, num_heads: int = 2, tf32: bool = False):
    """One optimisation step as trainer.py:365-391 does it (fp32, no AMP): mean cross-entropy over
    B*T frames (experiment.py:133), backward through the head only, one optimizer step.
    opt = {'name': 'sgd'|'adam'|'adamw', 'lr', 'weight_decay', ('momentum','dampening','nesterov') |
    ('beta1','beta2','eps')}; returns (loss, grads, new_sd, opt_state).  Optimizer arithmetic
    follows torch.optim.{SGD,Adam,AdamW} (instantiators.py:60-100).  tf32=True: the TCN convolutions' GEMMs
    (forward, dgrad, wgrad) see TF32-rounded operands, as the kernels' precision="tf32" mode does."""
    names = trainable_names(sd)
    P = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    buffers = {k: v.detach().clone() for k, v in sd.items() if k.startswith("bn.") and k not in P}
    with torch.enable_grad():
        logits = head_forward_train(P, buffers, {m: v.squeeze(1) if v.dim() == 4 else v for m, v in feats.items()},
                                    modalities, modal_dim, num_heads, seed, tf32=tf32)
        loss = F.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels.reshape(-1).long())
        grads = dict(zip(names, torch.autograd.grad(loss, [P[k] for k in names])))
    st = opt_state if opt_state is not None else {"step": 0, "m": {}, "v": {}}
    st["step"] += 1
    t = st["step"]
    new = {k: v.detach().clone() for k, v in sd.items()}
    new.update(buffers)
    lr, wd = opt["lr"], opt.get("weight_decay", 0.0)
    for k in names:
        p, g = P[k].detach(), grads[k]
        if opt["name"] == "sgd":
            mom, damp, nest = opt.get("momentum", 0.0), opt.get("dampening", 0.0), opt.get("nesterov", False)
            g = g + wd * p
            if mom != 0.0:
                buf = g.clone() if k not in st["m"] else st["m"][k] * mom + (1 - damp) * g
                st["m"][k] = buf
                g = g + mom * buf if nest else buf
            new[k] = p - lr * g
        else:
            b1, b2, eps = opt.get("beta1", 0.9), opt.get("beta2", 0.999), opt.get("eps", 1e-8)
            if opt["name"] == "adamw":
                p = p * (1 - lr * wd)
            else:
                g = g + wd * p
            m = st["m"].get(k, torch.zeros_like(p)) * b1 + (1 - b1) * g
            v = st["v"].get(k, torch.zeros_like(p)) * b2 + (1 - b2) * g * g
            st["m"][k], st["v"][k] = m, v
            denom = v.sqrt() / math.sqrt(1 - b2 ** t) + eps
            new[k] = p - (lr / (1 - b1 ** t)) * m / denom
    for k in list(new):            # keep the net.{0,4} aliases equal to conv1/conv2
        if ".net.0." in k:
            new[k] = new[k.replace(".net.0.", ".conv1.")]
        elif ".net.4." in k:
            new[k] = new[k.replace(".net.4.", ".conv2.")]
    return loss.detach(), grads, new, st


# --------------------------------------------------------------------------------------
# Eval-time input pipeline (base/dataset.py:503-510, base/transforms3D.py:15-144)
#   GroupNumpyToPILImage -> GroupScale(48) -> GroupCenterCrop(40) -> Stack ->
#   ToTorchFormatTensor (/255) -> GroupNormalize(mean .5, std .5)
# The resize is PIL's antialiased two-pass BILINEAR on uint8 (third-party: Pillow, un-pinned in the
# reference's environment; restated from its published algorithm -- libImaging/Resample.c:
# precompute_coeffs / normalize_coeffs_8bpc / ImagingResampleHorizontal_8bpc / Vertical_8bpc --
# and pinned bit-exactly against the Pillow installed in the build container, tests/golden/).
# --------------------------------------------------------------------------------------
PIL_PRECISION_BITS = 32 - 8 - 2


def pil_bilinear_coeffs(in_size: int, out_size: int):
    """precompute_coeffs + normalize_coeffs_8bpc for the triangle filter (support 1.0) over the
    whole axis.  Returns (bounds[out,2] = (xmin, count), kk[out, ksize] int32 fixed point)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.array([max(0.0, 1.0 - abs((x + xmin - center + 0.5) * ss)) for x in range(xmax)], dtype=np.float64)
        ww = 0.0
        for v in w:                       # sequential sum, as the C loop does
            ww += float(v)
        if ww != 0.0:
            w = w / ww
        for x in range(xmax):
            p = float(w[x]) * (1 << PIL_PRECISION_BITS)
            kk[xx, x] = int(-0.5 + p) if w[x] < 0 else int(0.5 + p)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pil_resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One 8bpc pass: out = clip8((2^(P-1) + sum pixel*kk) >> P) along ``axis`` of uint8 [H,W,C]."""
    bounds, kk = pil_bilinear_coeffs(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        xmin, cnt = bounds[xx]
        acc = (1 << (PIL_PRECISION_BITS - 1)) + np.tensordot(kk[xx, :cnt].astype(np.int64), src[xmin:xmin + cnt], axes=(0, 0))
        out[xx] = np.clip(acc >> PIL_PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """PIL.Image.resize((out_w, out_h), BILINEAR) on a uint8 [H,W,C] image: horizontal pass first,
    rounded to uint8, then the vertical pass (ImagingResample)."""
    tmp = _pil_resample_axis(img, out_w, 1) if out_w != img.shape[1] else img
    return _pil_resample_axis(tmp, out_h, 0) if out_h != img.shape[0] else tmp


def scaled_size(h: int, w: int, size: int):
    """torchvision.transforms.Resize(int): the smaller edge becomes ``size`` (GroupScale,
    base/transforms3D.py:103-117)."""
    if w <= h:
        return int(size * h / w), size
    return size, int(size * w / h)


def eval_transform(frames_u8: np.ndarray, size: int = 48, crop: int = 40) -> Tensor:
    """base/dataset.py:503-510 on uint8 [T,H,W,3] -> fp32 [T,3,crop,crop] in [-1,1]."""
    out = []
    for f in frames_u8:
        oh, ow = scaled_size(f.shape[0], f.shape[1], size)
        r = pil_resize_bilinear_u8(f, oh, ow)
        top, left = int(round((oh - crop) / 2.0)), int(round((ow - crop) / 2.0))       # torchvision CenterCrop
        out.append(r[top:top + crop, left:left + crop])
    x = torch.from_numpy(np.stack(out, 0)).permute(0, 3, 1, 2).contiguous().float().div(255)   # ToTorchFormatTensor
    return (x - 0.5) / 0.5                                                                      # Normalize(.5, .5)


def video_level_prediction(logits: np.ndarray, ignore_last_class: bool = False) -> Dict[str, int]:
    """format_trg_pred_video's three decision rules (metrics.py:118-142) for one video."""
    from collections import Counter
    lg = logits[:, :-1] if ignore_last_class else logits
    preds = np.argmax(lg, axis=1).flatten().tolist()
    vote = Counter(preds).most_common(1)[0][0]
    e = np.exp(lg)
    probs = e / np.sum(e, axis=1).reshape((-1, 1))
    return {"FRAMES_VOTE": int(vote), "FRAMES_AVG_LOGITS": int(np.argmax(lg.mean(axis=0))),
            "FRAMES_AVG_PROBS": int(np.argmax(probs.mean(axis=0)))}


# --------------------------------------------------------------------------------------
# Log-mel front end of the inline-VGGish modality
# (abaw5_pre_processing/base/vggish/mel_features.py:92-236, vggish_input.py:37-95,
#  vggish_params.py: 16 kHz, 25 ms periodic-Hann window, 10 ms hop, 512-point FFT, 64 mel bands
#  125-7500 Hz (HTK), log(mel + 0.01), examples of 96 frames)
# --------------------------------------------------------------------------------------
VGGISH_SR, VGGISH_WIN, VGGISH_HOP, VGGISH_FFT = 16000, 400, 160, 512
VGGISH_MEL_BINS, VGGISH_MEL_LO, VGGISH_MEL_HI, VGGISH_LOG_OFFSET = 64, 125.0, 7500.0, 0.01


def mel_matrix(num_mel_bins=VGGISH_MEL_BINS, num_spectrogram_bins=VGGISH_FFT // 2 + 1, sample_rate=VGGISH_SR,
               lower_hz=VGGISH_MEL_LO, upper_hz=VGGISH_MEL_HI) -> np.ndarray:
    """spectrogram_to_mel_matrix (mel_features.py:129-204): triangular bands, linear in HTK-mel
    (1127 ln(1 + f/700)), DC bin zeroed.  float64 [bins, mel]."""
    mel = lambda f: 1127.0 * np.log(1.0 + f / 700.0)
    bins_mel = mel(np.linspace(0.0, sample_rate / 2.0, num_spectrogram_bins))
    edges = np.linspace(mel(lower_hz), mel(upper_hz), num_mel_bins + 2)
    m = np.empty((num_spectrogram_bins, num_mel_bins))
    for i in range(num_mel_bins):
        lo, ce, up = edges[i:i + 3]
        m[:, i] = np.maximum(0.0, np.minimum((bins_mel - lo) / (ce - lo), (up - bins_mel) / (up - ce)))
    m[0, :] = 0.0
    return m


def log_mel_spectrogram(wave: np.ndarray) -> np.ndarray:
    """log_mel_spectrogram (mel_features.py:207-236) with the VGGish parameters: complete frames
    only (no padding), periodic Hann, |rfft(512)|, mel matrix, log(. + 0.01).  float64 [frames, 64]."""
    wave = np.asarray(wave, dtype=np.float64)
    n_frames = 1 + int(np.floor((wave.shape[0] - VGGISH_WIN) / VGGISH_HOP))
    idx = np.arange(VGGISH_WIN)[None, :] + VGGISH_HOP * np.arange(n_frames)[:, None]
    window = 0.5 - 0.5 * np.cos(2 * np.pi / VGGISH_WIN * np.arange(VGGISH_WIN))
    spec = np.abs(np.fft.rfft(wave[idx] * window, VGGISH_FFT))
    return np.log(spec @ mel_matrix() + VGGISH_LOG_OFFSET)


def example_starts(n_logmel_frames: int, window_sec: float, hop_sec: float):
    """my_frame's indexing (mel_features.py:21-46, vggish_input.py:70-79): examples of
    round(window_sec*100) frames, example i starts at Python-round(hop_sec*100 * i)."""
    win = int(round(window_sec * (1.0 / 0.010)))
    hop = hop_sec * (1.0 / 0.010)
    n = 1 + int(np.floor((n_logmel_frames - win) / hop))
    return [round(hop * i) for i in range(n)], win


def waveform_to_examples(wave: np.ndarray, window_sec: float = 0.96, hop_sec: float = 0.96) -> np.ndarray:
    """waveform_to_examples (vggish_input.py:37-82) for mono 16 kHz input: [n_examples, 96, 64]."""
    lm = log_mel_spectrogram(wave)
    starts, win = example_starts(lm.shape[0], window_sec, hop_sec)
    return np.stack([lm[s:s + win] for s in starts])



# --------------------------------------------------------------------------------------
# Alternative fusion heads: CAN (models/model.py:529-684), JMT / MT (:709-750, :895-1167)
# --------------------------------------------------------------------------------------
def _tcn_levels(sd: SD, prefix: str) -> int:
    n = 0
    while f"{prefix}network.{n}.conv1.weight_v" in sd:
        n += 1
    return n


def _encode_modalities(sd: SD, X: Dict[str, Tensor], modalities: Sequence[str]) -> Dict[str, Tensor]:
    """The part CAN.forward and JMT.forward share (model.py:655-676, :1138-1159): backbones, then
    TemporalConvNet + BatchNorm1d per modality.  Returns x[m]: [B, C_m, T] as the reference keeps it."""
    x = {}
    for m in modalities:
        if m == "video":
            B, T = X[m].shape[:2]
            f = ir50_forward(sd, X[m].reshape(B * T, *X[m].shape[2:]), "spatial.visual.backbone.").view(B, T, -1)
        elif m == "logmel":
            B, hh, T, ww = X[m].shape
            f = vggish_forward(sd, X[m].permute(0, 2, 3, 1).contiguous().view(-1, ww, hh), "spatial.audio.backbone.").view(B, T, -1)
        else:
            f = X[m].squeeze(1)
        h = tcn_forward(sd, f"temporal.{m}.", f.transpose(1, 2), _tcn_levels(sd, f"temporal.{m}."))
        x[m] = _bn_eval(sd, f"bn.{m}", h)
    return x


def _head_tail(sd: SD, c: Tensor, train_buffers: SD = None) -> Tensor:
    """fc1 -> BatchNorm1d over the feature axis -> F.leaky_relu -> fc2 (model.py:678-681, :1161-1164).  Eval statistics,
    or batch statistics (+ running-stat update in ``train_buffers``) in training mode."""
    c = F.linear(c, sd["fc1.weight"], sd["fc1.bias"]).transpose(1, 2)
    if train_buffers is None:
        c = _bn_eval(sd, "bn1", c).transpose(1, 2)
    else:
        c = F.batch_norm(c, train_buffers["bn1.running_mean"], train_buffers["bn1.running_var"], sd["bn1.weight"], sd["bn1.bias"],
                         True, BN_MOMENTUM, BN_EPS).transpose(1, 2)
        if "bn1.num_batches_tracked" in train_buffers:
            train_buffers["bn1.num_batches_tracked"] += 1
    return F.linear(F.leaky_relu(c, LEAKY_SLOPE), sd["fc2.weight"], sd["fc2.bias"])


def can_forward(sd: SD, X: Dict[str, Tensor], modalities: Sequence[str]) -> Tensor:
    """CAN.forward (model.py:651-684) with AttentionFusion (:552-568): per-modality Linear to 128,
    concat, softmax(Linear(concat)) as an element-wise gate.  Returns [B, T, output_dim]."""
    return can_fuse(sd, _encode_modalities(sd, X, modalities), modalities)


def can_fuse(sd: SD, x: Dict[str, Tensor], modalities: Sequence[str], train_buffers: SD = None) -> Tensor:
    """AttentionFusion + tail on the encoded modalities x[m]: [B, C_m, T]."""
    proj = [F.linear(x[m].transpose(1, 2), sd[f"fuse.attn.{i}.weight"], sd[f"fuse.attn.{i}.bias"])
            for i, m in enumerate(modalities)]
    cat = torch.cat(proj, -1)
    gate = torch.softmax(F.linear(cat, sd["fuse.weights.weight"], sd["fuse.weights.bias"]), dim=-1)
    return _head_tail(sd, gate * cat, train_buffers)


def mha1(sd: SD, p: str, q_in: Tensor, k_in: Tensor, v_in: Tensor) -> Tensor:
    """nn.MultiheadAttention(E, num_heads=1), sequence-first inputs [L, N, E] (model.py:731, :917-931)."""
    W, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    E = W.shape[1]
    q = F.linear(q_in, W[:E], b[:E]).transpose(0, 1)
    k = F.linear(k_in, W[E:2 * E], b[E:2 * E]).transpose(0, 1)
    v = F.linear(v_in, W[2 * E:], b[2 * E:]).transpose(0, 1)
    att = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(E), dim=-1)
    return F.linear((att @ v).transpose(0, 1), sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


def encoder_block(sd: SD, p: str, x: Tensor) -> Tensor:
    """TransformerEncoderBlock with one TransformerEncoderLayer (model.py:716-750): post-norm."""
    q = p + ".layers.0"
    x = x + mha1(sd, q + ".attention", x, x, x)
    x = F.layer_norm(x, (x.shape[-1],), sd[q + ".layer_norm1.weight"], sd[q + ".layer_norm1.bias"], LN_EPS)
    ff = F.linear(F.relu(F.linear(x, sd[q + ".feed_forward.0.weight"], sd[q + ".feed_forward.0.bias"])),
                  sd[q + ".feed_forward.2.weight"], sd[q + ".feed_forward.2.bias"])
    return F.layer_norm(x + ff, (x.shape[-1],), sd[q + ".layer_norm2.weight"], sd[q + ".layer_norm2.bias"], LN_EPS)


def jmt_forward(sd: SD, X: Dict[str, Tensor], modalities: Sequence[str], model_name: str = "JMT") -> Tensor:
    """JMT.forward (model.py:1134-1167) with JMTFusion (:933-979) or MTFusion (:1015-1048).
    Note the reference's final stage: the stacked cross-attention outputs [L, B, S, E] are viewed
    as [L*B, S, E] and fed to sequence-first attention modules, so the final encoder and
    self-attention attend over all L*B positions, with the S stack slots as the batch."""
    return jmt_fuse(sd, _encode_modalities(sd, X, modalities), model_name)


def jmt_fuse(sd: SD, x: Dict[str, Tensor], model_name: str = "JMT", train_buffers: SD = None) -> Tensor:
    """JMTFusion / MTFusion + tail on the encoded modalities x[m]: [B, C_m, T]."""
    f = "fuse."
    vis = x["video"].permute(2, 0, 1)
    aud = F.linear(x["vggish"].permute(2, 0, 1), sd[f + "augment_audio_feats_dim.weight"], sd[f + "augment_audio_feats_dim.bias"])
    ev, ea = encoder_block(sd, f + "visual_encoder", vis), encoder_block(sd, f + "audio_encoder", aud)
    if model_name == "JMT":
        jr = F.linear(torch.cat((vis, aud), dim=2), sd[f + "reduce_feats_dim.weight"], sd[f + "reduce_feats_dim.bias"])
        ej = encoder_block(sd, f + "jr_encoder", jr)
        stack = [mha1(sd, f + "CA_va", ev, ea, ea), mha1(sd, f + "CA_av", ea, ev, ev), mha1(sd, f + "CA_jrv", ej, ev, ev),
                 mha1(sd, f + "CA_vjr", ev, ej, ej), mha1(sd, f + "CA_jra", ej, ea, ea), mha1(sd, f + "CA_ajr", ea, ej, ej)]
    else:
        stack = [mha1(sd, f + "CA_va", ev, ea, ea), mha1(sd, f + "CA_av", ea, ev, ev)]
    st = torch.stack(stack, dim=2)
    L, B, S, E = st.shape
    st = st.view(-1, S, E)
    enc = encoder_block(sd, f + "final_encoder", st)
    out = mha1(sd, f + "final_self_attention", enc, enc, enc).view(L, B, S, E)[:, :, -1, :].permute(1, 0, 2)
    return _head_tail(sd, out, train_buffers)


ALT_TCN_DROPOUT = 0.2       # TemporalConvNet's default, which CAN / JMT do not override (model.py:592-596, :1079-1083)


def alt_head_trainable_names(sd: SD, name: str) -> List[str]:
    """Trainable tensors of CAN / JMT / MT that take part in forward: everything but spatial.*, buffers, the net.{0,4}
    aliases and the modules that are declared but never called (CAN's conv_c, MT's fuse.reduce_feats_dim: their .grad
    stays None and torch's optimizers skip them)."""
    skip = {"CAN": ("conv_c.",), "MT": ("fuse.reduce_feats_dim.",)}.get(name, ())     # MTFusion declares reduce_feats_dim, never calls it
    return [k for k in trainable_names(sd) if not k.startswith(skip)]


def alt_head_forward_train(name: str, P: SD, buffers: SD, feats: Dict[str, Tensor], modalities: Sequence[str], seed=None,
                           p_tcn: float = ALT_TCN_DROPOUT, tf32: bool = False) -> Tensor:
    """CAN / JMT / MT forward in TRAINING mode on pre-encoded features feats[m]: [B, T, D_m] (the frozen backbones
    stay in eval mode): TCN dropout + BatchNorm1d batch statistics (bn.<m> and bn1)."""
    enc = tcn_bn_train(P, buffers, feats, modalities, seed, p_tcn, tf32)
    x = {m: enc[m].transpose(1, 2) for m in modalities}
    if name == "CAN":
        return can_fuse(P, x, modalities, buffers)
    return jmt_fuse(P, x, name, buffers)


def alt_head_train_grads(name: str, sd: SD, feats: Dict[str, Tensor], labels: Tensor, modalities: Sequence[str], seed=None,
                         p_tcn: float = ALT_TCN_DROPOUT, tf32: bool = False):
    """(loss, grads, buffers, logits) of one mean-cross-entropy training step of CAN / JMT / MT (trainer.py:365-391)."""
    names = alt_head_trainable_names(sd, name)
    P = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    buffers = {k: v.detach().clone() for k, v in sd.items() if (k.startswith("bn.") or k.startswith("bn1.")) and k not in P}
    with torch.enable_grad():
        logits = alt_head_forward_train(name, P, buffers, feats, modalities, seed, p_tcn, tf32)
        loss = F.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels.reshape(-1).long())
        gl = torch.autograd.grad(loss, [P[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(P[k])) for k, g in zip(names, gl)}
    return loss.detach(), grads, buffers, logits.detach()



def attention_maps(sd: SD, prefix: str, feats: Dict[str, Tensor], modalities: Sequence[str], modal_dim: int = 32,
                   num_heads: int = 2) -> Tensor:
    """MultimodalTransformerEncoder.get_attention_maps (transformer.py:211-215): softmax(QK^T/sqrt(hd))
    over the modality tokens, [B, H, T, M, M]."""
    hd = modal_dim // num_heads
    a = prefix + "layers.self_attn."
    qs, ks = [], []
    for m in modalities:
        qkv = F.linear(feats[m], sd[f"{a}qkv_proj.{m}.weight"], sd[f"{a}qkv_proj.{m}.bias"])
        B, T, _ = qkv.shape
        qkv = qkv.view(B, T, num_heads, 3, hd)
        qs.append(qkv[:, :, :, 0]); ks.append(qkv[:, :, :, 1])
    Q, K = torch.stack(qs, 3), torch.stack(ks, 3)
    att = torch.softmax(torch.einsum("bthmd,bthnd->bthmn", Q, K) / math.sqrt(hd), dim=-1)
    return att.permute(0, 2, 1, 3, 4)
