"""Fusion-head training step on the B200 kernels (BASELINE config 4; trainer.py:365-391).

``HeadTrainer`` re-homes the trainable parameters of an LFAN mirror (everything but the frozen
``spatial.*`` backbones) into ONE flat fp32 buffer with a matching flat gradient buffer -- the
modules' ``nn.Parameter``s become views of it, so ``state_dict()`` / checkpoints keep the reference
layout -- and drives the CUDA plan (csrc/train.cu) through the C-ABI:

    forward (training mode: dropout + BatchNorm1d batch statistics)  ->  mean cross-entropy  ->
    backward  ->  [one NCCL all-reduce of the flat gradient buffer]  ->  fused SGD/Adam/AdamW.

Two ways in:
  * ``HeadTrainer.step(X, labels)`` -- the whole optimisation step in the kernels (what bench.py
    --workload train times);
  * ``LFAN.forward`` in training mode with grad enabled returns logits attached to autograd
    through ``_HeadFunction``; ``loss.backward()`` then runs the CUDA backward and fills ``.grad``,
    so the reference's loop (criterion, scaler.step(optimizer)) works unchanged.
PyTorch is used for memory, streams and the NCCL all-reduce only.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import _capi
from ._capi import HeadTrainSpec, check, lib

OPT_KINDS = {"sgd": 0, "adam": 1, "adamw": 2}


def head_parameters(model) -> List[tuple]:
    """(name, parameter) of the trainable head in ``named_parameters()`` order (the reference's
    optimizer sees exactly these: model.py freezes spatial.*)."""
    return [(k, p) for k, p in model.named_parameters() if not k.startswith("spatial.")]


def all_reduce_flat(flat: torch.Tensor, group=None) -> float:
    """ONE collective for the whole head: SUM all-reduce of the flat gradient bucket (any backend:
    nccl on the GPUs, gloo in the CPU tests).  Returns 1/world, the scale the optimizer kernel
    applies to turn the sum into the data-parallel mean; 1.0 when no process group is up."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world == 1:
        return 1.0
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


class HeadTrainer:
    def __init__(self, model, batch: int, length: Optional[int] = None, optimizer: Optional[dict] = None,
                 seed: int = 0, process_group=None, precision: str = "tf32"):
        """precision="tf32" (default): the GEMMs of forward, dgrad and wgrad run on TF32 tensor cores with
        fp32 accumulation -- what the reference's own GPU run does for its convolutions
        (torch.backends.cudnn.allow_tf32 defaults to True); "fp32": exact fp32 FMAs on CUDA cores
        (gradients within 2e-5 of CPU autograd, ~2.5x slower)."""
        _capi.require_gpu()
        if precision not in ("tf32", "fp32"):
            raise ValueError(precision)
        self.precision = precision
        self.model = model
        self.batch = int(batch)
        self.length = int(length or model.example_length)
        self.mods = list(model.modality)
        self.opt = dict(optimizer) if optimizer else None
        self.group = process_group
        self.base_seed = int(seed) & 0xFFFFFFFF
        self.calls = 0
        self.opt_step = 0
        named = head_parameters(model)
        self.device = named[0][1].device
        if self.device.type != "cuda":
            raise _capi.CerError("HeadTrainer needs the model on a CUDA device (no CPU fallback)")
        self.names = [k for k, _ in named]
        sizes = [p.numel() for _, p in named]
        self.count = sum(sizes)
        # flat buffers; every tensor starts on a 16-byte boundary
        offs, off = [], 0
        for n in sizes:
            offs.append(off)
            off += (n + 3) // 4 * 4
        self.flat_count = off
        self.params = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.grads = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.state_m = self.state_v = None
        self.offsets: Dict[str, int] = {}
        self._views: Dict[str, torch.Tensor] = {}
        self._gviews: Dict[str, torch.Tensor] = {}
        for (k, p), o, n in zip(named, offs, sizes):
            view = self.params[o:o + n].view(p.shape)
            view.copy_(p.data.float())
            p.data = view                                    # the module parameter now lives in the flat buffer
            self.offsets[k] = o
            self._views[k] = view
            self._gviews[k] = self.grads[o:o + n].view(p.shape)
        self._keep = []
        self._plans: Dict[tuple, tuple] = {}        # (B, T) -> (handle, workspace): all plans share the flat buffers
        self._gen = 0                               # bumped by every forward: the saved activations belong to the latest one
        self._build_spec()
        self._plan(self.batch, self.length)
        model.repack()

    # -- plan ---------------------------------------------------------------------------------
    def _pp(self, name):
        return self._views[name].data_ptr()

    def _gp(self, name):
        return self._gviews[name].data_ptr()

    def _build_spec(self):
        m = self.model
        s = HeadTrainSpec()
        s.n_modals = len(self.mods)
        s.kernel_size = m.kernel_size
        s.modal_dim, s.num_heads, s.n_out = m.modal_dim, m.num_heads, m.output_dim
        first = m.temporal[self.mods[0]].network[0]
        s.precision = 1 if self.precision == "tf32" else 0
        s.p_tcn = float(first.dropout1.p)
        s.p_fusion = float(m.fusion.layers.dropout.p)
        s.bn_momentum = float(m.bn[self.mods[0]].momentum)
        for mi, mod in enumerate(self.mods):
            tm = s.modal[mi]
            net = m.temporal[mod].network
            tm.in_dim, tm.n_blocks = m.embedding_dim[mod], len(net)
            if len(net) > _capi.CER_MAX_TCN_BLOCKS:
                raise ValueError(f"at most {_capi.CER_MAX_TCN_BLOCKS} TemporalBlocks per modality")
            for i, blk in enumerate(net):
                b = tm.blocks[i]
                p = f"temporal.{mod}.network.{i}."
                b.c_in, b.c_out, b.dilation = blk.conv1.in_channels, blk.conv1.out_channels, blk.dilation
                for conv, cn in ((b.conv1, "conv1"), (b.conv2, "conv2")):
                    conv.g, conv.v, conv.bias = self._pp(p + cn + ".weight_g"), self._pp(p + cn + ".weight_v"), self._pp(p + cn + ".bias")
                    conv.dg, conv.dv, conv.dbias = self._gp(p + cn + ".weight_g"), self._gp(p + cn + ".weight_v"), self._gp(p + cn + ".bias")
                if blk.downsample is not None:
                    b.wd, b.bd = self._pp(p + "downsample.weight"), self._pp(p + "downsample.bias")
                    b.dwd, b.dbd = self._gp(p + "downsample.weight"), self._gp(p + "downsample.bias")
            bn = m.bn[mod]
            for t in (bn.running_mean, bn.running_var):
                if not t.is_contiguous() or t.dtype != torch.float32:
                    raise ValueError("BatchNorm running statistics must be contiguous fp32")
            tm.bn_w, tm.bn_b = self._pp(f"bn.{mod}.weight"), self._pp(f"bn.{mod}.bias")
            tm.dbn_w, tm.dbn_b = self._gp(f"bn.{mod}.weight"), self._gp(f"bn.{mod}.bias")
            tm.bn_mean, tm.bn_var = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
            q = f"fusion.layers.self_attn.qkv_proj.{mod}."
            tm.wqkv, tm.bqkv, tm.dwqkv, tm.dbqkv = self._pp(q + "weight"), self._pp(q + "bias"), self._gp(q + "weight"), self._gp(q + "bias")
        a = "fusion.layers.self_attn.o_proj."
        n1 = "fusion.layers.norm1."
        s.wo, s.bo, s.dwo, s.dbo = self._pp(a + "weight"), self._pp(a + "bias"), self._gp(a + "weight"), self._gp(a + "bias")
        s.ln_g, s.ln_b, s.dln_g, s.dln_b = self._pp(n1 + "weight"), self._pp(n1 + "bias"), self._gp(n1 + "weight"), self._gp(n1 + "bias")
        s.wr, s.br, s.dwr, s.dbr = self._pp("regressor.weight"), self._pp("regressor.bias"), self._gp("regressor.weight"), self._gp("regressor.bias")
        s.grad_flat, s.grad_count = self.grads.data_ptr(), self.flat_count
        self._spec = s

    def _plan(self, batch: int, length: int):
        """The CUDA plan (workspace for saved activations + launch descriptors) for ``batch`` windows of
        ``length`` frames.  Plans are cached per shape -- the ragged last batch of an epoch gets its own --
        and all of them point at the same flat parameter / gradient buffers."""
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

    # -- pieces ---------------------------------------------------------------------------------
    def _feat_ptrs(self, feats: Dict[str, torch.Tensor]):
        f0 = feats[self.mods[0]]
        batch, length = int(f0.shape[0]), int(f0.shape[-2])
        rows = batch * length
        keep, ptrs = [], (C.c_void_p * len(self.mods))()
        for i, mod in enumerate(self.mods):
            f = feats[mod]
            if f.dim() == 4:
                f = f.squeeze(1)
            d = self.model.embedding_dim[mod]
            if tuple(f.shape) != (batch, length, d):
                raise ValueError(f"{mod}: expected [{batch},{length},{d}], got {tuple(f.shape)}")
            if f.device != self.device:
                raise ValueError("features must be on the trainer's CUDA device")
            f = f.float().contiguous().view(rows, d)
            keep.append(f)
            ptrs[i] = f.data_ptr()
        return keep, ptrs, batch, length

    def next_seed(self) -> int:
        s = (self.base_seed + self.calls * 0x632BE5AB) & 0xFFFFFFFF
        self.calls += 1
        return s

    def forward(self, feats: Dict[str, torch.Tensor], seed: Optional[int] = None) -> torch.Tensor:
        """Training-mode forward: logits [B, T, n_out]; keeps what backward needs."""
        keep, ptrs, batch, length = self._feat_ptrs(feats)
        h, _ = self._plan(batch, length)
        self._gen += 1
        self._last = (keep, ptrs, h, self._gen)
        seed = self.next_seed() if seed is None else int(seed) & 0xFFFFFFFF
        logits = torch.empty(batch, length, self.model.output_dim, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().cer_head_train_forward(h, ptrs, seed, logits.data_ptr(), _capi.current_stream_ptr()),
                  "cer_head_train_forward")
        for mod in self.mods:
            self.model.bn[mod].num_batches_tracked += 1
        return logits

    def cross_entropy(self, logits: torch.Tensor, labels: torch.Tensor, want_grad: bool = True):
        """Mean CE over B*T rows (trainer.py:372-381): returns (loss[1], dlogits or None)."""
        rows = logits.numel() // logits.shape[-1]
        lab = labels.reshape(rows).to(self.device, torch.int64).contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=self.device)
        dl = torch.empty_like(logits) if want_grad else None
        with torch.cuda.device(self.device):
            check(lib().cer_ce_loss(logits.contiguous().data_ptr(), lab.data_ptr(), rows, logits.shape[-1], loss.data_ptr(),
                                    None if dl is None else dl.data_ptr(), _capi.current_stream_ptr()), "cer_ce_loss")
        return loss, dl

    def backward(self, dlogits: torch.Tensor) -> None:
        """Fills the flat gradient buffer (overwrites) from d loss / d logits of the last forward."""
        keep, ptrs, h, _ = self._last
        dl = dlogits.float().contiguous()
        with torch.cuda.device(self.device):
            check(lib().cer_head_train_backward(h, ptrs, dl.data_ptr(), _capi.current_stream_ptr()),
                  "cer_head_train_backward")

    def grad(self, name: str) -> torch.Tensor:
        return self._gviews[name]

    def all_reduce_grads(self) -> float:
        """SUM all-reduce of the flat gradient buffer; returns the scale that makes it a mean."""
        return all_reduce_flat(self.grads, self.group)

    def apply_optimizer(self, grad_scale: float = 1.0) -> None:
        o = self.opt
        if o is None:
            raise ValueError("HeadTrainer was built without an optimizer config")
        kind = OPT_KINDS[o["name"]]
        if self.state_m is None:
            self.state_m = torch.zeros_like(self.params)
            self.state_v = torch.zeros_like(self.params) if kind > 0 else None
        self.opt_step += 1
        if kind == 0:
            b1, b2 = o.get("momentum", 0.0), o.get("dampening", 0.0)
        else:
            b1, b2 = o.get("beta1", 0.9), o.get("beta2", 0.999)
        with torch.cuda.device(self.device):
            check(lib().cer_optimizer_step(kind, self.params.data_ptr(), self.grads.data_ptr(), self.state_m.data_ptr(),
                                           None if self.state_v is None else self.state_v.data_ptr(), self.flat_count,
                                           o["lr"], o.get("weight_decay", 0.0), b1, b2, o.get("eps", 1e-8),
                                           int(bool(o.get("nesterov", False))), self.opt_step, grad_scale,
                                           _capi.current_stream_ptr()), "cer_optimizer_step")
        self.model.repack()                                   # inference engines hold packed copies of the old weights

    def step(self, X: Dict[str, torch.Tensor], labels: torch.Tensor, seed: Optional[int] = None,
             sync_grads: bool = True) -> torch.Tensor:
        """One optimisation step (trainer.py:365-391): returns the loss as a 1-element device tensor.
        ``sync_grads=False`` skips the gradient all-reduce (a collective: every rank of the group must
        call it) -- for a step that only some ranks of an initialised process group take."""
        logits = self.forward(X, seed)
        loss, dl = self.cross_entropy(logits, labels)
        self.backward(dl)
        scale = self.all_reduce_grads() if sync_grads else 1.0
        self.apply_optimizer(scale)
        return loss


class _HeadFunction(torch.autograd.Function):
    """Autograd bridge: forward/backward of the head run in the CUDA plan; the parameters are
    passed as inputs only so that autograd routes their gradients."""

    @staticmethod
    def forward(ctx, trainer: HeadTrainer, feats: dict, *params):
        ctx.trainer = trainer
        out = trainer.forward(feats)
        ctx.gen = trainer._gen
        return out

    @staticmethod
    def backward(ctx, dlogits):
        tr = ctx.trainer
        if ctx.gen != tr._gen:
            # the plan keeps ONE set of saved activations: a later training forward has overwritten the ones
            # this graph needs -- fail instead of returning gradients of the wrong batch
            raise RuntimeError("LFAN training forward was called again before backward() of an earlier output: "
                               "run backward (or drop the earlier output) before the next training forward")
        tr.backward(dlogits)
        return (None, None) + tuple(tr.grad(k).clone() for k in tr.names)


def forward_with_grad(model, feats: Dict[str, torch.Tensor]) -> torch.Tensor:
    """LFAN.forward's training branch: logits [B, T, n_out] attached to autograd."""
    any_f = feats[model.modality[0]]
    B, T = any_f.shape[0], any_f.shape[-2]
    tr = model.__dict__.get("_trainer")
    if tr is None:
        # other (B, T) shapes add a plan to the same trainer; model.train_precision ("tf32" | "fp32") picks the GEMMs
        tr = HeadTrainer(model, B, T, precision=getattr(model, "train_precision", "tf32"))
        model.__dict__["_trainer"] = tr
    params = [p for _, p in head_parameters(model)]
    return _HeadFunction.apply(tr, feats, *params)
