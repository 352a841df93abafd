"""GPU parity of TRAINING the alternative heads CAN / JMT / MT (experiment.py:317-347 trains them through the same
loop as LFAN): the backward building blocks against PyTorch autograd, and one whole training step against the
REFERENCE's own autograd (tests/golden/heads_train.pt: reference modules, Dropout p = 0) and against the oracle with
this repo's dropout masks.  Exact-fp32 mode carries the tight bars; TF32 mode is held to the TF32 deviation of these
models (see tests/test_gpu_train.py)."""
import os
import warnings

import pytest
import torch
import torch.nn.functional as F

from feature_vs_text_compound_emotion_b200 import synthetic
from oracle import lfan_oracle as O

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")
BS = {"visual_state_dict": "res50_ir_0.887", "audio_state_dict": "vggish"}
DIMS = {"video": 512, "vggish": 128, "bert": 768}


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _head(name, dev, p_drop=None):
    from feature_vs_text_compound_emotion_b200.models.model import CAN, JMT
    mods = ["video", "vggish", "bert"] if name == "CAN" else ["video", "vggish"]
    vsd = synthetic.visual_backbone_state_dict(0)
    if name == "CAN":
        m = CAN(task="CLASSIFICATION", modalities=mods, tcn_settings=synthetic.TCN_SETTINGS, backbone_settings=BS, output_dim=7,
                root_dir="", device=dev, visual_state_dict=vsd)
        sd = synthetic.can_state_dict(0, mods)
    else:
        m = JMT(task="CLASSIFICATION", modalities=mods, tcn_settings=synthetic.TCN_SETTINGS, backbone_settings=BS, output_dim=7,
                root_dir="", device=dev, model_name=name, visual_state_dict=vsd)
        sd = synthetic.jmt_state_dict(0, mods, model_name=name)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).train()
    if p_drop is not None:
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = p_drop
    return m, sd, mods


def _grads_close(tr, ref_grads, rel, what):
    """Every gradient tensor within rel * ||ref_k|| + 5e-5 * max_k ||ref|| element-wise (L-inf).  The second term is for
    tensors whose true gradient is ZERO -- a bias in front of a batch-statistics BatchNorm (fc1.bias: the reference's own
    value is 8.9e-6 of round-off), the discarded attention branches: what is left there is summation noise of the real
    gradients' magnitude."""
    scale = max(float(g.double().norm()) for g in ref_grads.values())
    worst = 0.0
    for k, g in ref_grads.items():
        mine = tr.grad(k).cpu().double()
        ref = g.double()
        if mine.shape != ref.shape:                                  # a 1/97 sample of a large tensor
            mine = mine.flatten()[::97]
        tol = rel * float(ref.norm()) + 5e-5 * scale
        err = float((mine - ref).abs().max())
        worst = max(worst, err / max(tol, 1e-30))
        assert err <= tol, (what, k, err, tol)
    return worst


def test_backward_building_blocks_vs_autograd():
    """cer_linear_backward, cer_act_backward, cer_softmax_gate_backward, cer_add_layernorm_backward,
    cer_bn1d_train_forward / _backward, cer_sdpa_train_forward / cer_sdpa_backward (self- and cross-attention on
    column slices of packed projections, ragged lengths) and cer_add_inplace against torch autograd on the CPU."""
    import ctypes as C
    from feature_vs_text_compound_emotion_b200 import _capi, engine as E
    dev = _dev()
    lib, st = _capi.lib(), _capi.current_stream_ptr
    g = torch.Generator().manual_seed(41)
    R, I, N = 333, 96, 200
    x = torch.randn(R, I, generator=g)
    w = (torch.randn(N, I, generator=g) * 0.1).requires_grad_(True)
    b = torch.randn(N, generator=g).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    dy = torch.randn(R, N, generator=g)
    with torch.enable_grad():
        (F.linear(xr, w, b) * dy).sum().backward()
    xd, wd, dyd = x.to(dev), w.detach().to(dev), dy.to(dev)
    dx, dw, db = torch.empty(R, I, device=dev), torch.zeros(N, I, device=dev), torch.zeros(N, device=dev)
    _capi.check(lib.cer_linear_backward(xd.data_ptr(), R, I, I, wd.data_ptr(), dyd.data_ptr(), N, N, dx.data_ptr(), I, 0, dw.data_ptr(),
                                        db.data_ptr(), st()))
    assert (dx.cpu() - xr.grad).abs().max() < 1e-4 and (dw.cpu() - w.grad).abs().max() < 2e-4 and (db.cpu() - b.grad).abs().max() < 2e-4
    _capi.check(lib.cer_linear_backward(xd.data_ptr(), R, I, I, wd.data_ptr(), dyd.data_ptr(), N, N, dx.data_ptr(), I, 1, None, None, st()))
    assert (dx.cpu() - 2 * xr.grad).abs().max() < 2e-4                                   # dx_accumulate
    # activations
    y = torch.randn(1000, generator=g)
    for act, fn in ((1, F.leaky_relu), (2, F.relu)):
        yr = y.clone().requires_grad_(True)
        with torch.enable_grad():
            out = fn(yr)
            out.backward(torch.ones_like(out) * 3)
        d, od, dd = torch.empty(1000, device=dev), out.detach().to(dev), (torch.ones(1000) * 3).to(dev)     # keep the operands alive
        _capi.check(lib.cer_act_backward(act, od.data_ptr(), dd.data_ptr(), 1000, d.data_ptr(), st()))
        assert (d.cpu() - yr.grad).abs().max() < 1e-6
    # softmax gate
    gate, feat, dyg = (torch.randn(77, 384, generator=g) for _ in range(3))
    gr, fr = gate.clone().requires_grad_(True), feat.clone().requires_grad_(True)
    with torch.enable_grad():
        ((torch.softmax(gr, -1) * fr) * dyg).sum().backward()
    dgt, dft = torch.empty(77, 384, device=dev), torch.empty(77, 384, device=dev)
    gd, fd, dygd = gate.to(dev), feat.to(dev), dyg.to(dev)
    _capi.check(lib.cer_softmax_gate_backward(gd.data_ptr(), fd.data_ptr(), dygd.data_ptr(), 77, 384, dgt.data_ptr(), dft.data_ptr(), st()))
    assert (dgt.cpu() - gr.grad).abs().max() < 1e-6 and (dft.cpu() - fr.grad).abs().max() < 1e-6
    # residual + LayerNorm
    xs, rs, dyl = (torch.randn(301, 128, generator=g) for _ in range(3))
    gam, bet = (torch.rand(128, generator=g) + 0.5).requires_grad_(True), torch.randn(128, generator=g).requires_grad_(True)
    xsr, rsr = xs.clone().requires_grad_(True), rs.clone().requires_grad_(True)
    with torch.enable_grad():
        (F.layer_norm(xsr + rsr, (128,), gam, bet, 1e-5) * dyl).sum().backward()
    dxl, dga, dbe = torch.empty(301, 128, device=dev), torch.zeros(128, device=dev), torch.zeros(128, device=dev)
    xsd, rsd, gamd, dyld = xs.to(dev), rs.to(dev), gam.detach().to(dev), dyl.to(dev)
    _capi.check(lib.cer_add_layernorm_backward(xsd.data_ptr(), rsd.data_ptr(), 301, 128, gamd.data_ptr(), 1e-5, dyld.data_ptr(), dxl.data_ptr(),
                                               dga.data_ptr(), dbe.data_ptr(), st()))
    assert (dxl.cpu() - xsr.grad).abs().max() < 1e-5 and torch.equal(xsr.grad, rsr.grad)
    assert (dga.cpu() - gam.grad).abs().max() < 1e-4 and (dbe.cpu() - bet.grad).abs().max() < 1e-4
    # BatchNorm1d, training mode
    xb, dyb = torch.randn(500, 70, generator=g) * 2 + 1, torch.randn(500, 70, generator=g)
    wb, bb = (torch.rand(70, generator=g) + 0.5).requires_grad_(True), torch.randn(70, generator=g).requires_grad_(True)
    rm, rv = torch.zeros(70), torch.ones(70)
    xbr = xb.clone().requires_grad_(True)
    with torch.enable_grad():
        yb = F.batch_norm(xbr, rm, rv, wb, bb, True, 0.1, 1e-5)
        (yb * dyb).sum().backward()
    yd, mean, inv = torch.empty(500, 70, device=dev), torch.empty(70, device=dev), torch.empty(70, device=dev)
    rmd, rvd = torch.zeros(70, device=dev), torch.ones(70, device=dev)
    xbd, wbd, bbd, dybd = xb.to(dev), wb.detach().to(dev), bb.detach().to(dev), dyb.to(dev)
    _capi.check(lib.cer_bn1d_train_forward(xbd.data_ptr(), 500, 70, wbd.data_ptr(), bbd.data_ptr(), yd.data_ptr(),
                                           mean.data_ptr(), inv.data_ptr(), rmd.data_ptr(), rvd.data_ptr(), 0.1, st()))
    assert (yd.cpu() - yb.detach()).abs().max() < 1e-5 and (rmd.cpu() - rm).abs().max() < 1e-6 and (rvd.cpu() - rv).abs().max() < 1e-5
    dxb, dwb, dbb = torch.empty(500, 70, device=dev), torch.empty(70, device=dev), torch.empty(70, device=dev)
    _capi.check(lib.cer_bn1d_train_backward(dybd.data_ptr(), xbd.data_ptr(), 500, 70, wbd.data_ptr(), mean.data_ptr(),
                                            inv.data_ptr(), dxb.data_ptr(), dwb.data_ptr(), dbb.data_ptr(), st()))
    assert (dxb.cpu() - xbr.grad).abs().max() < 1e-5 and (dwb.cpu() - wb.grad).abs().max() < 1e-4 and (dbb.cpu() - bb.grad).abs().max() < 1e-4
    # attention with saved probabilities
    for (e, batch, lq, lk, self_attn) in ((128, 2, 300, 300, True), (128, 3, 37, 91, False), (64, 1, 130, 130, True)):
        qkv = torch.randn(batch * max(lq, lk), 3 * e, generator=g)
        do = torch.randn(batch * lq, e, generator=g)
        q, k, v = (qkv[:batch * n_, c0:c0 + e].clone().requires_grad_(True) for n_, c0 in ((lq, 0), (lk, e), (lk, 2 * e)))
        with torch.enable_grad():
            want = F.scaled_dot_product_attention(q.view(batch, lq, e), k.view(batch, lk, e), v.view(batch, lk, e)).reshape(-1, e)
            (want * do).sum().backward()
        qd = qkv.to(dev)
        qs, ks, vs = qd[:batch * lq, :e], qd[:batch * lk, e:2 * e], qd[:batch * lk, 2 * e:]
        out = torch.empty(batch * lq, e, device=dev)
        probs, scratch = torch.empty(batch, lq, lk, device=dev), torch.empty(batch, lq, lk, device=dev)
        _capi.check(lib.cer_sdpa_train_forward(qs.data_ptr(), 3 * e, ks.data_ptr(), 3 * e, vs.data_ptr(), 3 * e, batch, lq, lk, e, out.data_ptr(), e,
                                               probs.data_ptr(), st()))
        assert (out.cpu() - want.detach()).abs().max() < 2e-5
        dqkv = torch.zeros(batch * max(lq, lk), 3 * e, device=dev)
        dod = do.to(dev)
        _capi.check(lib.cer_sdpa_backward(qs.data_ptr(), 3 * e, ks.data_ptr(), 3 * e, vs.data_ptr(), 3 * e, probs.data_ptr(), dod.data_ptr(), e,
                                          batch, lq, lk, e, dqkv[:, :e].data_ptr(), 3 * e, dqkv[:, e:2 * e].data_ptr(), 3 * e,
                                          dqkv[:, 2 * e:].data_ptr(), 3 * e, scratch.data_ptr(), st()))
        got = dqkv.cpu()
        assert (got[:batch * lq, :e] - q.grad).abs().max() < 5e-5, (e, batch, lq, lk)
        assert (got[:batch * lk, e:2 * e] - k.grad).abs().max() < 5e-5 and (got[:batch * lk, 2 * e:] - v.grad).abs().max() < 5e-5
        del self_attn
    a_, b_ = torch.randn(1000, generator=g), torch.randn(1000, generator=g)
    assert torch.equal(E.add_(a_.to(dev), b_.to(dev)).cpu(), a_ + b_)
    del C


@pytest.mark.parametrize("name", ["CAN", "JMT", "MT"])
def test_training_step_vs_reference_golden(golden_dir, name):
    """One training step of the exact-fp32 mode against the REFERENCE's autograd (reference modules in train mode on the
    same 512-d embeddings, Dropout p = 0): loss within 2e-5, logits within 2e-4, BatchNorm running statistics within 1e-5,
    the never-called modules keep no gradient, and every gradient element within
      * CAN: 5e-4 of its tensor's norm (measured: worst 8.6e-5, median 5.5e-7);
      * JMT / MT: 6e-3.  These gradients are ill-conditioned in fp32: the REFERENCE's own fp32 result is 0.8e-3 (JMT) /
        2.2e-3 (MT) of a tensor norm away from an fp64 evaluation of the same model (oracle in float64) on the bias
        gradients of the final encoder and self-attention -- column sums with heavy cancellation behind LayerNorm and a
        batch-statistics BatchNorm over 80 rows -- with a median of 2e-5 / 5e-5 over all tensors.  The kernels sit at
        the same level against the reference: worst 1.4e-3 / 3.1e-3, median 3.5e-5 / 2.9e-4.
    """
    from feature_vs_text_compound_emotion_b200.heads_training import AltHeadTrainer
    dev = _dev()
    g = torch.load(os.path.join(golden_dir, "heads_train.pt"))[name]
    m, sd, mods = _head(name, dev, p_drop=0.0)
    assert mods == g["modalities"]
    B, T = g["B"], g["T"]
    gen = torch.Generator().manual_seed(g["seed"])
    feats = {k: torch.randn(B, T, DIMS[k], generator=gen) for k in mods}
    labels = torch.randint(0, 7, (B, T, 1), generator=gen)
    tr = AltHeadTrainer(m, B, T, precision="fp32")
    logits = tr.forward({k: v.to(dev) for k, v in feats.items()})
    assert (logits.cpu() - g["logits"]).abs().max().item() < 2e-4
    loss, dl = tr.cross_entropy(logits, labels.to(dev))
    assert abs(loss.item() - g["loss"]) < 2e-5
    tr.backward(dl)
    ref = dict(g["grad_small"])
    ref.update(g["grad_sample"])
    assert set(ref) == set(tr.names) and not (set(g["none_grad"]) & set(tr.names))
    _grads_close(tr, ref, 5e-4 if name == "CAN" else 6e-3, name)
    for k, gn in g["grad_norm"].items():
        assert abs(float(tr.grad(k).double().norm()) - gn) <= 5e-3 * gn + 2e-5 * max(g["grad_norm"].values()), k
    cur = m.state_dict()
    for k, v in g["bn"].items():
        assert (cur[k].cpu() - v).abs().max().item() < 1e-5, k


@pytest.mark.parametrize("name", ["CAN", "JMT"])
def test_training_step_with_dropout_vs_oracle_and_tf32(name):
    """Dropout ON in the TCN stacks (p = 0.2, TemporalConvNet's default) against the oracle with the kernels' mask
    hash, exact mode; then the default TF32 mode on the same inputs within the TF32 deviation."""
    from feature_vs_text_compound_emotion_b200.heads_training import AltHeadTrainer
    dev = _dev()
    gen = torch.Generator().manual_seed(77)
    seed = 0xBEEF
    out = {}
    for prec in ("fp32", "tf32"):
        m, sd, mods = _head(name, dev)
        feats = {k: torch.randn(2, 60, DIMS[k], generator=torch.Generator().manual_seed(78)) for k in mods}
        labels = torch.randint(0, 7, (2, 60, 1), generator=torch.Generator().manual_seed(79))
        tr = AltHeadTrainer(m, 2, 60, precision=prec)
        logits = tr.forward({k: v.to(dev) for k, v in feats.items()}, seed=seed)
        loss, dl = tr.cross_entropy(logits, labels.to(dev))
        tr.backward(dl)
        out[prec] = (logits.cpu(), loss.item(), tr)
    ref_loss, grads, _, ref_logits = O.alt_head_train_grads(name, sd, feats, labels, mods, seed=seed)
    lf, lossf, trf = out["fp32"]
    assert (lf - ref_logits).abs().max().item() < 2e-4 and abs(lossf - float(ref_loss)) < 2e-5
    _grads_close(trf, grads, 5e-4 if name == "CAN" else 6e-3, name + " fp32 vs oracle")
    lt, losst, trt = out["tf32"]
    assert (lt - lf).abs().max().item() < 2e-2 and not torch.equal(lt, lf) and abs(losst - lossf) < 5e-3
    scale = max(float(v.double().norm()) for v in grads.values())
    for k, v in grads.items():
        a, b = trt.grad(k).cpu().double().flatten(), v.double().flatten()
        assert float((a - b).norm()) <= 0.3 * max(float(b.norm()), 1e-2 * scale), k
    del gen


def test_module_training_loop_and_eval_after_step():
    """The reference's loop shape on the drop-in module: model.train(); out = model(X) from pixels (frozen IR-50);
    loss.backward(); torch optimizer step; then model.eval() sees the updated weights.  Also AltHeadTrainer.step with
    AdamW lowers the loss on a repeated batch."""
    from feature_vs_text_compound_emotion_b200.heads_training import AltHeadTrainer
    dev = _dev()
    m, sd, mods = _head("CAN", dev, p_drop=0.0)
    m.train_precision = "fp32"
    T = 40
    vid = synthetic.frames(T, seed=91).view(1, T, 3, 40, 40)
    f = synthetic.feature_windows(1, T, seed=92, modalities=["vggish", "bert"])
    labels = torch.randint(0, 7, (1, T), generator=torch.Generator().manual_seed(93))
    X = lambda: {"video": vid.to(dev), "vggish": f["vggish"].to(dev), "bert": f["bert"].to(dev)}
    m.eval()
    with torch.no_grad():
        before = m(X()).clone()
    m.train()
    opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=0.5)
    with torch.enable_grad():
        out = m(X())
        assert out.requires_grad and out.shape == (1, T, 7)
        loss = F.cross_entropy(out.view(T, 7), labels.view(T).to(dev))
        loss.backward()
    assert all(p.grad is None for p in m.spatial.parameters()) and m.conv_c.weight.grad is None
    assert m.fc2.weight.grad is not None and float(m.fc2.weight.grad.abs().max()) > 0
    opt.step()
    m.eval()
    with torch.no_grad():
        after = m(X())
    assert (after - before).abs().max().item() > 1e-3
    # ... and they are the oracle's logits on the UPDATED weights (embeddings from the frozen kernels both times)
    with torch.no_grad():
        emb = m.spatial["visual"](vid.view(T, 3, 40, 40).to(dev)).cpu().view(1, T, 512)
    new_sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    enc = {"video": emb, "vggish": f["vggish"].squeeze(1), "bert": f["bert"].squeeze(1)}
    xs = {k: O._bn_eval(new_sd, f"bn.{k}", O.tcn_forward(new_sd, f"temporal.{k}.", enc[k].transpose(1, 2),
                                                         O._tcn_levels(new_sd, f"temporal.{k}."))) for k in mods}
    want = O.can_fuse(new_sd, xs, mods)
    assert (after.cpu() - want).abs().max().item() <= 2e-2
    m2, _, mods2 = _head("JMT", dev, p_drop=0.0)
    tr = AltHeadTrainer(m2, 2, 60, optimizer={"name": "adamw", "lr": 1e-3, "weight_decay": 1e-4}, seed=3)
    feats = {k: torch.randn(2, 60, DIMS[k], generator=torch.Generator().manual_seed(94)).to(dev) for k in mods2}
    y = torch.randint(0, 7, (2, 60, 1), generator=torch.Generator().manual_seed(95)).to(dev)
    losses = [tr.step(feats, y).item() for _ in range(40)]
    assert losses[-1] < 0.9 * losses[0] and max(losses[-5:]) < min(losses[:3]), losses
