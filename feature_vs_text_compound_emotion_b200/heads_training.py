"""Training step of the alternative heads CAN / JMT / MT on the B200 kernels.

The reference trains these heads through the same loop as LFAN (experiment.py:317-347, trainer.py:365-391:
mean cross-entropy over B*T frames, frozen backbones).  ``AltHeadTrainer`` mirrors ``training.HeadTrainer``:
the trainable parameters move into ONE flat fp32 buffer (the modules' ``nn.Parameter``s become views of it, so
``state_dict()`` keeps the reference layout) with a matching flat gradient buffer, and one step is

    forward (training mode: TCN dropout, BatchNorm1d batch statistics for bn.<m> and bn1)  ->  cer_ce_loss  ->
    backward  ->  [one all-reduce of the flat gradient buffer]  ->  cer_optimizer_step.

The TemporalConvNet + BatchNorm1d stacks run in the CUDA training plan LFAN uses (cer_head_train_tcn_*, TF32 tensor
cores by default); everything after them is composed here from the C-ABI building blocks -- cer_linear_forward /
_backward, cer_softmax_gate(_backward), cer_add_layernorm(_backward), cer_sdpa_train_forward / cer_sdpa_backward,
cer_bn1d_train_*, cer_act_backward -- each backward written out by hand below, mirroring heads.py's forward.
PyTorch owns memory and the NCCL all-reduce; no torch op computes anything.

What takes part (checked against the reference's own autograd, tests/golden/heads_train.pt): every parameter except
the frozen ``spatial.*`` and the modules that are declared but never called (CAN.conv_c, MTFusion.reduce_feats_dim:
their ``.grad`` stays None in the reference and torch's optimizers skip them).  Branches of JMT / MT whose outputs
are discarded (all stacked cross-attentions but the last, model.py:975) get exactly-zero gradients, as in the
reference, and are therefore still subject to weight decay.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _capi
from . import engine as E
from ._capi import HeadTrainSpec, check, lib
from .training import OPT_KINDS, all_reduce_flat

UNUSED = {"CAN": ("conv_c.",), "JMT": (), "MT": ("fuse.reduce_feats_dim.",)}


def _ptr(t):
    return None if t is None else t.data_ptr()


class AltHeadTrainer:
    def __init__(self, model, batch: int, length: int, optimizer: Optional[dict] = None, seed: int = 0, process_group=None,
                 precision: str = "tf32"):
        from .heads import CAN, JMT
        _capi.require_gpu()
        if isinstance(model, CAN):
            self.kind = "CAN"
        elif isinstance(model, JMT):
            self.kind = model.model_name
        else:
            raise TypeError("AltHeadTrainer trains heads.CAN / heads.JMT (LFAN: training.HeadTrainer)")
        if precision not in ("tf32", "fp32"):
            raise ValueError(precision)
        self.model, self.precision = model, precision
        self.batch, self.length = int(batch), int(length)
        self.mods = list(model.modalities)
        self.opt = dict(optimizer) if optimizer else None
        self.group = process_group
        self.base_seed = int(seed) & 0xFFFFFFFF
        self.calls = self.opt_step = 0
        named = [(k, p) for k, p in model.named_parameters()
                 if not k.startswith("spatial.") and not k.startswith(UNUSED[self.kind])]
        self.device = named[0][1].device
        if self.device.type != "cuda":
            raise _capi.CerError("AltHeadTrainer needs the model on a CUDA device (no CPU fallback)")
        self.names = [k for k, _ in named]
        offs, off = [], 0
        for _, p in named:
            offs.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.flat_count = off
        self.params = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.grads = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.state_m = self.state_v = None
        self._views: Dict[str, torch.Tensor] = {}
        self._gviews: Dict[str, torch.Tensor] = {}
        for (k, p), o in zip(named, offs):
            view = self.params[o:o + p.numel()].view(p.shape)
            view.copy_(p.data.float())
            p.data = view
            self._views[k] = view
            self._gviews[k] = self.grads[o:o + p.numel()].view(p.shape)
        self._plans: Dict[tuple, tuple] = {}
        self._build_spec()
        self._plan(self.batch, self.length)
        model.repack()

    # -- TCN + BatchNorm1d plan (shared with LFAN's training plan) ---------------------------------
    def P(self, name):
        return self._views[name]

    def G(self, name):
        return self._gviews[name]

    def _build_spec(self):
        m = self.model
        s = HeadTrainSpec()
        s.n_modals = len(self.mods)
        if s.n_modals > _capi.CER_MAX_MODALS:
            raise ValueError("too many modalities")
        first = m.temporal[self.mods[0]].network[0]
        s.kernel_size = first.conv1.kernel_size[0]
        s.modal_dim, s.num_heads, s.n_out = 32, 2, 1              # fusion part of the plan is never run (tcn-only calls)
        s.precision = 1 if self.precision == "tf32" else 0
        s.p_tcn = float(first.dropout1.p)
        s.p_fusion = 0.0
        s.bn_momentum = float(m.bn[self.mods[0]].momentum)
        self.c_last = {}
        for mi, mod in enumerate(self.mods):
            tm = s.modal[mi]
            net = m.temporal[mod].network
            if len(net) > _capi.CER_MAX_TCN_BLOCKS:
                raise ValueError(f"at most {_capi.CER_MAX_TCN_BLOCKS} TemporalBlocks per modality")
            tm.in_dim, tm.n_blocks = net[0].conv1.in_channels, len(net)
            for i, blk in enumerate(net):
                b = tm.blocks[i]
                p = f"temporal.{mod}.network.{i}."
                if blk.conv1.kernel_size[0] != s.kernel_size:
                    raise ValueError("all TemporalBlocks must share one kernel size")
                b.c_in, b.c_out, b.dilation = blk.conv1.in_channels, blk.conv1.out_channels, blk.dilation
                for conv, cn in ((b.conv1, "conv1"), (b.conv2, "conv2")):
                    conv.g, conv.v, conv.bias = (self.P(p + cn + x).data_ptr() for x in (".weight_g", ".weight_v", ".bias"))
                    conv.dg, conv.dv, conv.dbias = (self.G(p + cn + x).data_ptr() for x in (".weight_g", ".weight_v", ".bias"))
                if blk.downsample is not None:
                    b.wd, b.bd = self.P(p + "downsample.weight").data_ptr(), self.P(p + "downsample.bias").data_ptr()
                    b.dwd, b.dbd = self.G(p + "downsample.weight").data_ptr(), self.G(p + "downsample.bias").data_ptr()
            bn = m.bn[mod]
            tm.bn_w, tm.bn_b = self.P(f"bn.{mod}.weight").data_ptr(), self.P(f"bn.{mod}.bias").data_ptr()
            tm.dbn_w, tm.dbn_b = self.G(f"bn.{mod}.weight").data_ptr(), self.G(f"bn.{mod}.bias").data_ptr()
            tm.bn_mean, tm.bn_var = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
            self.c_last[mod] = net[-1].conv1.out_channels
        s.grad_flat, s.grad_count = self.grads.data_ptr(), self.flat_count
        self._spec = s

    def _plan(self, batch: int, length: int):
        key = (int(batch), int(length))
        p = self._plans.get(key)
        if p is None:
            with torch.cuda.device(self.device):
                nbytes = lib().cer_head_train_workspace_bytes(C.byref(self._spec), key[0], key[1])
                if nbytes == 0:
                    raise _capi.CerError("cer_head_train_workspace_bytes rejected the spec: " + (lib().cer_last_error() or b"").decode())
                ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
                h = C.c_void_p()
                check(lib().cer_head_train_create(C.byref(h), C.byref(self._spec), key[0], key[1], ws.data_ptr(), nbytes),
                      "cer_head_train_create")
            p = self._plans[key] = (h, ws)
        return p

    def __del__(self):
        for h, _ in getattr(self, "_plans", {}).values():
            try:
                lib().cer_head_train_destroy(h)
            except Exception:
                pass
        self._plans = {}

    # -- building blocks: forward returns what backward needs ------------------------------------------------
    def _stream(self):
        return _capi.current_stream_ptr(self.device)

    def _lin_bwd(self, x, wname, dy, dx=None, accumulate=False, w=None, dw=None, db=None, bias=True):
        """dy [rows, out] (any row pitch) through y = x W^T + b: returns dx (allocated unless given)."""
        W = self.P(wname + ".weight") if w is None else w
        dW = self.G(wname + ".weight") if dw is None else dw
        dB = (self.G(wname + ".bias") if db is None else db) if bias else None
        if W.dim() == 3:                       # Conv1d(k=1) weights [out][in][1]
            W, dW = W[:, :, 0], dW[:, :, 0]
        rows, out_dim = dy.shape
        in_dim = W.shape[1]
        if dx is None:
            dx = torch.empty(rows, in_dim, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().cer_linear_backward(x.data_ptr(), rows, in_dim, x.stride(0), W.data_ptr(), dy.data_ptr(), out_dim, dy.stride(0),
                                            dx.data_ptr(), dx.stride(0), int(accumulate), dW.data_ptr(), _ptr(dB), self._stream()),
                  "cer_linear_backward")
        return dx

    def _act_bwd(self, act, y, dy):
        dx = torch.empty_like(dy)
        with torch.cuda.device(self.device):
            check(lib().cer_act_backward(E.ACT[act], y.data_ptr(), dy.contiguous().data_ptr(), y.numel(), dx.data_ptr(), self._stream()),
                  "cer_act_backward")
        return dx

    def _ln_bwd(self, x, res, lnname, dy, eps):
        dx = torch.empty_like(x)
        with torch.cuda.device(self.device):
            check(lib().cer_add_layernorm_backward(x.data_ptr(), _ptr(res), x.shape[0], x.shape[1], self.P(lnname + ".weight").data_ptr(), eps,
                                                   dy.contiguous().data_ptr(), dx.data_ptr(), self.G(lnname + ".weight").data_ptr(),
                                                   self.G(lnname + ".bias").data_ptr(), self._stream()), "cer_add_layernorm_backward")
        return dx

    def _mha_fwd(self, name, q_in, kv_in, batch, len_q, len_k):
        """nn.MultiheadAttention(E, 1)(q_in, kv_in, kv_in) on batch-major rows; keeps the probabilities."""
        W, b = self.P(name + ".in_proj_weight"), self.P(name + ".in_proj_bias")
        e = W.shape[1]
        if q_in is kv_in:
            qkv = E.linear(q_in, W, b)
            q, k, v = qkv[:, :e], qkv[:, e:2 * e], qkv[:, 2 * e:]
        else:
            q = E.linear(q_in, W[:e], b[:e])
            kv = E.linear(kv_in, W[e:], b[e:])
            k, v = kv[:, :e], kv[:, e:]
        att = torch.empty(batch * len_q, e, dtype=torch.float32, device=self.device)
        probs = torch.empty(batch, len_q, len_k, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().cer_sdpa_train_forward(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), batch, len_q,
                                               len_k, e, att.data_ptr(), e, probs.data_ptr(), self._stream()), "cer_sdpa_train_forward")
        out = E.linear(att, self.P(name + ".out_proj.weight"), self.P(name + ".out_proj.bias"))
        return out, (name, q_in, kv_in, q, k, v, att, probs, batch, len_q, len_k)

    def _mha_bwd(self, saved, d_out):
        """Returns (d q_in, d kv_in); for self-attention the two are one tensor (already summed)."""
        name, q_in, kv_in, q, k, v, att, probs, batch, len_q, len_k = saved
        e = att.shape[1]
        d_att = self._lin_bwd(att, name + ".out_proj", d_out)
        W, dW, dB = self.P(name + ".in_proj_weight"), self.G(name + ".in_proj_weight"), self.G(name + ".in_proj_bias")
        scratch = torch.empty_like(probs)
        if q_in is kv_in:
            dqkv = torch.zeros(batch * len_q, 3 * e, dtype=torch.float32, device=self.device)
            dq, dk, dv = dqkv[:, :e], dqkv[:, e:2 * e], dqkv[:, 2 * e:]
        else:
            dq = torch.empty(batch * len_q, e, dtype=torch.float32, device=self.device)
            dkv = torch.zeros(batch * len_k, 2 * e, dtype=torch.float32, device=self.device)
            dk, dv = dkv[:, :e], dkv[:, e:]
        with torch.cuda.device(self.device):
            check(lib().cer_sdpa_backward(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), probs.data_ptr(),
                                          d_att.data_ptr(), e, batch, len_q, len_k, e, dq.data_ptr(), dq.stride(0), dk.data_ptr(),
                                          dk.stride(0), dv.data_ptr(), dv.stride(0), scratch.data_ptr(), self._stream()),
                  "cer_sdpa_backward")
        if q_in is kv_in:
            d_in = self._lin_bwd(q_in, None, dqkv, w=W, dw=dW, db=dB)
            return d_in, d_in
        d_q = self._lin_bwd(q_in, None, dq, w=W[:e], dw=dW[:e], db=dB[:e])
        d_kv = self._lin_bwd(kv_in, None, dkv, w=W[e:], dw=dW[e:], db=dB[e:])
        return d_q, d_kv

    def _enc_fwd(self, name, x, batch, length):
        """TransformerEncoderBlock with one post-norm layer (model.py:716-750) on batch-major rows."""
        p = name + ".layers.0"
        ln1, ln2 = getattr(self._mod(p), "layer_norm1"), getattr(self._mod(p), "layer_norm2")
        a, sa = self._mha_fwd(p + ".attention", x, x, batch, length, length)
        x1 = E.add_layernorm(x, a, self.P(p + ".layer_norm1.weight"), self.P(p + ".layer_norm1.bias"), ln1.eps)
        h = E.linear(x1, self.P(p + ".feed_forward.0.weight"), self.P(p + ".feed_forward.0.bias"), "relu")
        ff = E.linear(h, self.P(p + ".feed_forward.2.weight"), self.P(p + ".feed_forward.2.bias"))
        x2 = E.add_layernorm(x1, ff, self.P(p + ".layer_norm2.weight"), self.P(p + ".layer_norm2.bias"), ln2.eps)
        return x2, (p, x, a, sa, x1, h, ff, ln1.eps, ln2.eps)

    def _enc_bwd(self, saved, d_x2):
        p, x, a, sa, x1, h, ff, eps1, eps2 = saved
        d_s2 = self._ln_bwd(x1, ff, p + ".layer_norm2", d_x2, eps2)                 # d(x1 + ff): both branches
        d_h = self._act_bwd("relu", h, self._lin_bwd(h, p + ".feed_forward.2", d_s2))
        d_x1 = self._lin_bwd(x1, p + ".feed_forward.0", d_h, dx=d_s2, accumulate=True)   # + the residual path
        d_s1 = self._ln_bwd(x, a, p + ".layer_norm1", d_x1, eps1)                   # d(x + a)
        d_in, _ = self._mha_bwd(sa, d_s1)                                           # through the attention branch ...
        return E.add_(d_in, d_s1)                                                   # ... plus the residual branch

    def _mod(self, dotted):
        m = self.model
        for part in dotted.split("."):
            m = m[int(part)] if part.isdigit() else getattr(m, part)
        return m

    # -- the step ----------------------------------------------------------------------------------------------
    def next_seed(self) -> int:
        s = (self.base_seed + self.calls * 0x632BE5AB) & 0xFFFFFFFF
        self.calls += 1
        return s

    def _tcn_forward(self, feats, seed):
        f0 = feats[self.mods[0]]
        B, T = int(f0.shape[0]), int(f0.shape[-2])
        R = B * T
        h, _ = self._plan(B, T)
        keep, ptrs = [], (C.c_void_p * len(self.mods))()
        zs, zptrs = {}, (C.c_void_p * len(self.mods))()
        for i, mod in enumerate(self.mods):
            f = feats[mod]
            if f.dim() == 4:
                f = f.squeeze(1)
            if f.shape[0] != B or f.shape[1] != T or f.device != self.device:
                raise ValueError(f"{mod}: expected [{B},{T},D] on {self.device}, got {tuple(f.shape)} on {f.device}")
            f = f.float().contiguous().view(R, -1)
            keep.append(f)
            ptrs[i] = f.data_ptr()
            zs[mod] = torch.empty(R, self.c_last[mod], dtype=torch.float32, device=self.device)
            zptrs[i] = zs[mod].data_ptr()
        with torch.cuda.device(self.device):
            check(lib().cer_head_train_tcn_forward(h, ptrs, seed, zptrs, self._stream()), "cer_head_train_tcn_forward")
        for mod in self.mods:
            self.model.bn[mod].num_batches_tracked += 1
        return zs, (h, keep, ptrs, B, T)

    def _tail_fwd(self, c):
        m = self.model
        h1 = E.linear(c, self.P("fc1.weight"), self.P("fc1.bias"))
        n = h1.shape[1]
        h1n = torch.empty_like(h1)
        mean, inv = torch.empty(n, device=self.device), torch.empty(n, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().cer_bn1d_train_forward(h1.data_ptr(), h1.shape[0], n, self.P("bn1.weight").data_ptr(), self.P("bn1.bias").data_ptr(),
                                               h1n.data_ptr(), mean.data_ptr(), inv.data_ptr(), m.bn1.running_mean.data_ptr(),
                                               m.bn1.running_var.data_ptr(), float(m.bn1.momentum), self._stream()), "cer_bn1d_train_forward")
            a = torch.empty_like(h1n)
            check(lib().cer_leaky_relu_forward(h1n.data_ptr(), h1n.numel(), a.data_ptr(), self._stream()), "cer_leaky_relu_forward")
        m.bn1.num_batches_tracked += 1
        out = E.linear(a, self.P("fc2.weight"), self.P("fc2.bias"))
        return out, (c, h1, mean, inv, a)

    def _tail_bwd(self, saved, d_out):
        c, h1, mean, inv, a = saved
        d_h1n = self._act_bwd("leaky_relu", a, self._lin_bwd(a, "fc2", d_out))
        d_h1 = torch.empty_like(h1)
        with torch.cuda.device(self.device):
            check(lib().cer_bn1d_train_backward(d_h1n.data_ptr(), h1.data_ptr(), h1.shape[0], h1.shape[1], self.P("bn1.weight").data_ptr(),
                                                mean.data_ptr(), inv.data_ptr(), d_h1.data_ptr(), self.G("bn1.weight").data_ptr(),
                                                self.G("bn1.bias").data_ptr(), self._stream()), "cer_bn1d_train_backward")
        return self._lin_bwd(c, "fc1", d_h1)

    def forward(self, feats: Dict[str, torch.Tensor], seed: Optional[int] = None) -> torch.Tensor:
        """feats[m]: [B, T, D_m] pre-encoded features (visual = 512-d IR-50 embeddings) -> logits [B, T, n_out]."""
        seed = self.next_seed() if seed is None else int(seed) & 0xFFFFFFFF
        z, tcn_saved = self._tcn_forward(feats, seed)
        B, T = tcn_saved[3], tcn_saved[4]
        R = B * T
        if self.kind == "CAN":
            n = len(self.mods)
            cat = torch.empty(R, 128 * n, dtype=torch.float32, device=self.device)
            for i, mod in enumerate(self.mods):
                E.linear(z[mod], self.P(f"fuse.attn.{i}.weight"), self.P(f"fuse.attn.{i}.bias"), out=cat[:, 128 * i:128 * (i + 1)])
            gate = E.linear(cat, self.P("fuse.weights.weight"), self.P("fuse.weights.bias"))
            fused = E.softmax_gate(gate, cat)
            out, tail = self._tail_fwd(fused)
            self._saved = ("CAN", tcn_saved, z, cat, gate, fused, tail)
        else:
            f = "fuse."
            vis = z["video"]
            aud = E.linear(z["vggish"], self.P(f + "augment_audio_feats_dim.weight"), self.P(f + "augment_audio_feats_dim.bias"))
            ea, s_ea = self._enc_fwd(f + "audio_encoder", aud, B, T)
            if self.kind == "JMT":
                va = torch.cat((vis, aud), dim=1)
                jr = E.linear(va, self.P(f + "reduce_feats_dim.weight"), self.P(f + "reduce_feats_dim.bias"))
                eo, s_eo = self._enc_fwd(f + "jr_encoder", jr, B, T)
                last, s_ca = self._mha_fwd(f + "CA_ajr", ea, eo, B, T, T)
            else:
                va = None
                eo, s_eo = self._enc_fwd(f + "visual_encoder", vis, B, T)
                last, s_ca = self._mha_fwd(f + "CA_av", ea, eo, B, T, T)
            enc, s_fe = self._enc_fwd(f + "final_encoder", last, 1, R)
            fin, s_fa = self._mha_fwd(f + "final_self_attention", enc, enc, 1, R, R)
            out, tail = self._tail_fwd(fin)
            self._saved = (self.kind, tcn_saved, z, vis, aud, va, s_ea, s_eo, s_ca, s_fe, s_fa, tail)
        return out.view(B, T, -1)

    def backward(self, dlogits: torch.Tensor) -> None:
        """Fills the flat gradient buffer from d loss / d logits of the last forward."""
        sv = self._saved
        self.grads.zero_()
        tcn_saved, z = sv[1], sv[2]
        h, keep, ptrs, B, T = tcn_saved
        R = B * T
        dl = dlogits.float().contiguous().view(R, -1)
        dz = {}
        if sv[0] == "CAN":
            _, _, _, cat, gate, fused, tail = sv
            d_fused = self._tail_bwd(tail, dl)
            d_gate, d_cat = torch.empty_like(gate), torch.empty_like(cat)
            with torch.cuda.device(self.device):
                check(lib().cer_softmax_gate_backward(gate.data_ptr(), cat.data_ptr(), d_fused.data_ptr(), R, cat.shape[1], d_gate.data_ptr(),
                                                      d_cat.data_ptr(), self._stream()), "cer_softmax_gate_backward")
            self._lin_bwd(cat, "fuse.weights", d_gate, dx=d_cat, accumulate=True)
            for i, mod in enumerate(self.mods):
                dz[mod] = self._lin_bwd(z[mod], f"fuse.attn.{i}", d_cat[:, 128 * i:128 * (i + 1)])
        else:
            kind, _, _, vis, aud, va, s_ea, s_eo, s_ca, s_fe, s_fa, tail = sv
            f = "fuse."
            d_fin = self._tail_bwd(tail, dl)
            d_enc, _ = self._mha_bwd(s_fa, d_fin)
            d_last = self._enc_bwd(s_fe, d_enc)
            d_ea, d_eo = self._mha_bwd(s_ca, d_last)
            d_aud = self._enc_bwd(s_ea, d_ea)
            d_o = self._enc_bwd(s_eo, d_eo)
            if kind == "JMT":
                d_va = self._lin_bwd(va, f + "reduce_feats_dim", d_o)            # [R, 256] = d(vis | aud)
                dz["video"] = d_va[:, :128].contiguous()
                d_aud = E.add_(d_aud, d_va[:, 128:].contiguous())
            else:
                dz["video"] = d_o
            dz["vggish"] = self._lin_bwd(z["vggish"], f + "augment_audio_feats_dim", d_aud)
        dzp = (C.c_void_p * len(self.mods))()
        keep_dz = []
        for i, mod in enumerate(self.mods):
            t = dz[mod].contiguous()
            keep_dz.append(t)
            dzp[i] = t.data_ptr()
        with torch.cuda.device(self.device):
            check(lib().cer_head_train_tcn_backward(h, ptrs, dzp, self._stream()), "cer_head_train_tcn_backward")

    def cross_entropy(self, logits: torch.Tensor, labels: torch.Tensor, want_grad: bool = True):
        rows = logits.numel() // logits.shape[-1]
        lab = labels.reshape(rows).to(self.device, torch.int64).contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=self.device)
        dl = torch.empty_like(logits) if want_grad else None
        with torch.cuda.device(self.device):
            check(lib().cer_ce_loss(logits.contiguous().data_ptr(), lab.data_ptr(), rows, logits.shape[-1], loss.data_ptr(), _ptr(dl),
                                    self._stream()), "cer_ce_loss")
        return loss, dl

    def grad(self, name: str) -> torch.Tensor:
        return self._gviews[name]

    def apply_optimizer(self, grad_scale: float = 1.0) -> None:
        o = self.opt
        if o is None:
            raise ValueError("AltHeadTrainer was built without an optimizer config")
        kind = OPT_KINDS[o["name"]]
        if self.state_m is None:
            self.state_m = torch.zeros_like(self.params)
            self.state_v = torch.zeros_like(self.params) if kind > 0 else None
        self.opt_step += 1
        b1, b2 = (o.get("momentum", 0.0), o.get("dampening", 0.0)) if kind == 0 else (o.get("beta1", 0.9), o.get("beta2", 0.999))
        with torch.cuda.device(self.device):
            check(lib().cer_optimizer_step(kind, self.params.data_ptr(), self.grads.data_ptr(), self.state_m.data_ptr(), _ptr(self.state_v),
                                           self.flat_count, o["lr"], o.get("weight_decay", 0.0), b1, b2, o.get("eps", 1e-8),
                                           int(bool(o.get("nesterov", False))), self.opt_step, grad_scale, self._stream()),
                  "cer_optimizer_step")
        self.model.repack()

    def step(self, feats: Dict[str, torch.Tensor], labels: torch.Tensor, seed: Optional[int] = None,
             sync_grads: bool = True) -> torch.Tensor:
        """One optimisation step (trainer.py:365-391): returns the loss as a 1-element device tensor.
        ``sync_grads=False`` skips the gradient all-reduce (a collective: every rank of the group must
        call it) -- for a step that only some ranks of an initialised process group take."""
        logits = self.forward(feats, seed)
        loss, dl = self.cross_entropy(logits, labels)
        self.backward(dl)
        self.apply_optimizer(all_reduce_flat(self.grads, self.group) if sync_grads else 1.0)
        return loss


class _AltHeadFunction(torch.autograd.Function):
    """Autograd bridge (as training._HeadFunction): forward / backward run in the kernels, the parameters are inputs
    only so that autograd routes their gradients into ``.grad``."""

    @staticmethod
    def forward(ctx, trainer: AltHeadTrainer, feats: dict, *params):
        ctx.trainer = trainer
        out = trainer.forward(feats)
        ctx.gen = trainer.calls
        return out

    @staticmethod
    def backward(ctx, dlogits):
        tr = ctx.trainer
        if ctx.gen != tr.calls:
            raise RuntimeError("the head's training forward was called again before backward() of an earlier output: "
                               "run backward (or drop the earlier output) before the next training forward")
        tr.backward(dlogits)
        return (None, None) + tuple(tr.grad(k).clone() for k in tr.names)


def forward_with_grad(model, feats: Dict[str, torch.Tensor]) -> torch.Tensor:
    """CAN / JMT ``forward`` in training mode with grad enabled: logits [B, T, n_out] attached to autograd, so the
    reference's loop (criterion, loss.backward(), any torch optimizer) works unchanged."""
    f0 = feats[list(model.modalities)[0]]
    tr = model.__dict__.get("_trainer")
    if tr is None:
        tr = AltHeadTrainer(model, int(f0.shape[0]), int(f0.shape[-2]), precision=getattr(model, "train_precision", "tf32"))
        model.__dict__["_trainer"] = tr
    params = [dict(model.named_parameters())[k] for k in tr.names]
    return _AltHeadFunction.apply(tr, feats, *params)
