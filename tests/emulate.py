"""Pure-torch emulation of what the CUDA kernels compute from the PACKED weights (tests only).

Used on the CPU to prove the packing algebra (BN folding, border-class bias table, fused
projection shortcut, NHWC FC permutation, weight-norm, tap-major TCN weights) before any GPU
time is spent, and to predict the bf16 error budget.  `round_act=True` rounds every stored
activation to bf16 exactly where the kernels do.
"""
import torch
import torch.nn.functional as F


def _r(x, on):
    return x.to(torch.bfloat16).float() if on else x


def border_class_map(H, W):
    r = torch.ones(H, dtype=torch.long); r[0] = 0; r[-1] = 2
    c = torch.ones(W, dtype=torch.long); c[0] = 0; c[-1] = 2
    return r.view(-1, 1) * 3 + c.view(1, -1)          # [H, W]


def ir50_packed(pk, x, round_act=True, upto=None):
    """x: [N,3,H,W] fp32 -> (emb [N,512], dict of unit outputs NHWC)."""
    N, _, H, W = x.shape
    w = pk["stem_w"].view(3, 3, 3, 64).permute(3, 2, 0, 1)           # [(r,s,ci)][co] -> [co,ci,r,s]
    h = F.conv2d(x, w, pk["stem_bias"], 1, 1)
    a = pk["stem_alpha"].view(1, -1, 1, 1)
    h = _r(torch.where(h >= 0, h, h * a), round_act)
    taps = {-1: h.permute(0, 2, 3, 1).contiguous()}
    for i, u in enumerate(pk["units"]):
        cin, depth, stride = u["cin"], u["depth"], u["stride"]
        w1 = u["w1"].float().view(depth, 3, 3, cin).permute(0, 3, 1, 2)
        t = F.conv2d(h, w1, None, 1, 1)
        cls = border_class_map(h.shape[2], h.shape[3])
        t = t + u["bias1"][cls].permute(2, 0, 1).unsqueeze(0)        # [9,depth] indexed by class map
        al = u["alpha"].view(1, -1, 1, 1)
        t = _r(torch.where(t >= 0, t, t * al), round_act)
        w2 = u["w2"].float()
        w2m = w2[:, :9 * depth].view(depth, 3, 3, depth).permute(0, 3, 1, 2)
        y = F.conv2d(t, w2m, None, stride, 1)
        if u["has_proj"]:
            y = y + F.conv2d(h, w2[:, 9 * depth:].view(depth, cin, 1, 1), None, stride)
        else:
            y = y + h
        h = _r(y + u["bias2"].view(1, -1, 1, 1), round_act)
        taps[i] = h.permute(0, 2, 3, 1).contiguous()
        if upto is not None and i == upto:
            return None, taps
    flat = h.permute(0, 2, 3, 1).reshape(N, -1)
    e = F.linear(flat, pk["fc_w"].float(), pk["fc_bias"])
    return e / e.norm(dim=1, keepdim=True), taps


def tcn_packed(blocks, x):
    """x: [B,T,Cin] -> [B,T,Cout], from tap-major packed weights."""
    for blk in blocks:
        k, d = blk["kernel_size"], blk["dilation"]
        def cconv(inp, w, b):
            B, T, _ = inp.shape
            out = b.view(1, 1, -1).expand(B, T, -1).clone()
            for j in range(k):
                sh = (k - 1 - j) * d
                shifted = F.pad(inp, (0, 0, sh, 0))[:, :T]
                out = out + shifted @ w[j]
            return out
        h = F.leaky_relu(cconv(x, blk["w1"], blk["b1"]), 0.01)
        h = F.leaky_relu(cconv(h, blk["w2"], blk["b2"]), 0.01)
        res = x if blk["wd"] is None else x @ blk["wd"] + blk["bd"]
        x = F.leaky_relu(h + res, 0.01)
        if blk["post_scale"] is not None:
            x = x * blk["post_scale"] + blk["post_shift"]
    return x


def fusion_packed(fw, feats):
    """feats: list of [R, D_m] -> (logits [R,n_out], fused [R,E])."""
    M, md, H = fw["n_modals"], fw["modal_dim"], fw["num_heads"]
    hd = md // H
    qkv = [f @ w + b for f, w, b in zip(feats, fw["wqkv"], fw["bqkv"])]          # [R, 3*md]
    R = feats[0].shape[0]
    q = torch.stack([t.view(R, H, 3, hd)[:, :, 0] for t in qkv], dim=2)           # [R,H,M,hd]
    k = torch.stack([t.view(R, H, 3, hd)[:, :, 1] for t in qkv], dim=2)
    v = torch.stack([t.view(R, H, 3, hd)[:, :, 2] for t in qkv], dim=2)
    att = torch.softmax(q @ k.transpose(-1, -2) / hd ** 0.5, dim=-1)
    vals = (att @ v + v).reshape(R, H * M * hd)
    o = vals @ fw["wo"] + fw["bo"]
    fused = F.layer_norm(o, (o.shape[-1],), fw["ln_g"], fw["ln_b"], 1e-5)
    logits = torch.cat([feats[0], fused], dim=-1) @ fw["wr"] + fw["br"]
    return logits, fused
