"""GPU parity of the fusion-head training step (BASELINE config 4) -- forward in training mode,
cross-entropy, backward, optimizer -- against the reference-generated golden (Dropout p = 0) and
against the oracle with this repo's dropout masks.  Tolerances: fp32 kernels vs fp32 CPU autograd,
gradients within 2e-5 of each tensor's norm (8x what was measured, see _grad_close), loss within 2e-5."""
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
MODS = ["cnn_res50", "vggish", "bert"]


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _lfan(mods, dev, seed=0, length=300, p_drop=None, precision="fp32"):
    from feature_vs_text_compound_emotion_b200.models.model import LFAN
    m = LFAN(backbone_settings=BS, output_dim=7, task="CLASSIFICATION", modality=mods, kernel_size=5,
             example_length=length, tcn_channel=synthetic.TCN_CHANNELS, modal_dim=32, num_heads=2, root_dir="",
             device=dev)
    m.init()
    m.load_state_dict(synthetic.lfan_state_dict(seed, mods), strict=True)
    m = m.to(dev).train()
    m.train_precision = precision            # what LFAN.forward's training branch hands to HeadTrainer
    if p_drop is not None:
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = p_drop
    return m


def _grad_close(mine, ref, norm=None, rel=2e-5):
    """Gradient parity of the exact-fp32 training kernels against fp32 CPU autograd.
    Measured on B200 (gpurun_out/r02_grad_stats.json, 888 tensor comparisons of this file): NO element
    further than 2.5e-6 * ||ref|| from the reference, relative L2 error <= 3.8e-6.  The bounds are set
    8x above that: every element within 2e-5 * ||ref||, except that a LeakyReLU kink may move a few
    (a pre-activation within fp32 round-off of zero gets slope 1 or 0.01 depending on the summation
    order, here as in MKL -- none was observed, so the allowance is 2 elements or 0.02 % of the tensor),
    and 1e-3 in relative L2."""
    mine, ref = mine.double().flatten(), ref.double().flatten()
    norm = float(ref.norm()) if norm is None else norm
    err = (mine - ref).abs()
    n_bad = int((err > rel * norm + 1e-7).sum())
    assert n_bad <= max(2, mine.numel() // 5000), (n_bad, mine.numel(), err.max().item())
    assert float(err.norm()) <= 1e-3 * norm + 1e-6, (float(err.norm()), norm)
    _STATS.append({"numel": mine.numel(), "n_bad": n_bad, "rel_l2": float(err.norm()) / max(norm, 1e-30),
                   "max_rel": float(err.max()) / max(norm, 1e-30)})
    return n_bad


_STATS = []


@pytest.fixture(scope="module", autouse=True)
def _dump_grad_stats():
    """CER_GRAD_STATS=<path>: write what _grad_close observed (how many elements per tensor were out of
    tolerance, relative L2 error) so the bounds above can be set from measurements."""
    yield
    path = os.environ.get("CER_GRAD_STATS")
    if path and _STATS:
        import json
        with open(path, "w") as f:
            json.dump(_STATS, f)


def _check_grads(tr, grads, rel=2e-5):
    kinks = 0
    for k, g in grads.items():
        try:
            kinks += _grad_close(tr.grad(k).cpu(), g, rel=rel) > 0
        except AssertionError as e:
            raise AssertionError(f"{k}: {e}") from None
    assert kinks <= 3, f"{kinks} tensors with out-of-tolerance elements: more than LeakyReLU kinks explain"
    return kinks


def test_two_sgd_steps_vs_reference_golden(golden_dir):
    from feature_vs_text_compound_emotion_b200.training import HeadTrainer
    dev = _dev()
    g = torch.load(os.path.join(golden_dir, "train_b2.pt"))
    mods = g["modalities"]
    m = _lfan(mods, dev, g["weights_seed"], p_drop=0.0)
    tr = HeadTrainer(m, 2, 300, optimizer=g["opt"], precision="fp32")
    X = {k: v.to(dev) for k, v in synthetic.feature_windows(2, 300, seed=g["x_seed"], modalities=mods).items()}
    labels = torch.randint(0, 7, (2, 300, 1), generator=torch.Generator().manual_seed(g["label_seed"])).float().to(dev)
    for step in g["steps"]:
        loss = tr.step(X, labels)
        assert abs(loss.item() - step["loss"]) < 2e-5
        for k, gn in step["grad_norm"].items():
            mine = tr.grad(k).cpu()
            ref = step["grad_small"][k] if k in step["grad_small"] else step["grad_sample"][k]
            got = mine if k in step["grad_small"] else mine.flatten()[::997]
            _grad_close(got, ref, norm=gn)
            assert abs(float(mine.double().norm()) - gn) <= 2e-2 * gn + 1e-7, k
        sd = m.state_dict()
        for k, v in step["bn"].items():
            assert (sd[k].cpu().float() - v.float()).abs().max().item() < 1e-5, k
        for k, v in step["param_sample"].items():
            d = (sd[k].cpu().flatten()[::997] - v).abs()
            assert int((d > 1e-5).sum()) <= 1 and d.max().item() < 1e-4, k


def test_dropout_forward_backward_vs_oracle():
    """Dropout ON (p = 0.1 as model.py:471,482): the kernels' counter-hash masks are restated by
    the oracle, so logits, loss and every gradient are compared element by element."""
    from feature_vs_text_compound_emotion_b200.training import HeadTrainer
    dev = _dev()
    m = _lfan(MODS, dev, seed=3)
    tr = HeadTrainer(m, 2, 300, precision="fp32")
    sd = synthetic.lfan_state_dict(3, MODS)
    X = synthetic.feature_windows(2, 300, seed=21, modalities=MODS)
    labels = torch.randint(0, 7, (2, 300, 1), generator=torch.Generator().manual_seed(22)).float()
    seed = 0xC0FFEE
    logits = tr.forward({k: v.to(dev) for k, v in X.items()}, seed=seed)
    loss, dl = tr.cross_entropy(logits, labels.to(dev))
    tr.backward(dl)
    P = {k: sd[k] for k in O.trainable_names(sd)}
    buffers = {k: v.clone() for k, v in sd.items() if k.startswith("bn.") and k not in P}
    want = O.head_forward_train(P, buffers, {k: v.squeeze(1) for k, v in X.items()}, MODS, seed=seed)
    assert (logits.cpu() - want).abs().max().item() < 2e-4
    ref_loss, grads, _, _ = O.train_step(sd, X, labels, MODS, {"name": "sgd", "lr": 0.0}, None, seed=seed)
    assert abs(loss.item() - float(ref_loss)) < 2e-5
    _check_grads(tr, grads)
    # a different seed draws different masks
    logits2 = tr.forward({k: v.to(dev) for k, v in X.items()}, seed=seed + 1)
    assert (logits2 - logits).abs().max().item() > 1e-3


@pytest.mark.parametrize("name,cfg", [
    ("adamw", {"name": "adamw", "lr": 1e-3, "weight_decay": 1e-2}),
    ("adam", {"name": "adam", "lr": 1e-3, "weight_decay": 1e-4}),
    ("sgd_plain", {"name": "sgd", "lr": 5e-2, "momentum": 0.0, "weight_decay": 0.0}),
])
def test_optimizers_vs_oracle(name, cfg):
    from feature_vs_text_compound_emotion_b200.training import HeadTrainer
    dev = _dev()
    mods = ["vggish", "bert"]          # a two-modality head: E = 64, leader width 32
    m = _lfan(mods, dev, seed=1, length=120, p_drop=0.0)
    tr = HeadTrainer(m, 3, 120, optimizer=cfg, precision="fp32")
    sd = synthetic.lfan_state_dict(1, mods)
    X = synthetic.feature_windows(3, 120, seed=31, modalities=mods)
    labels = torch.randint(0, 7, (3, 120, 1), generator=torch.Generator().manual_seed(32)).float()
    st = None
    for it in range(2):
        loss = tr.step({k: v.to(dev) for k, v in X.items()}, labels.to(dev))
        before = sd
        ref_loss, grads, sd, st = O.train_step(sd, X, labels, mods, cfg, st)
        if it == 0:
            assert abs(loss.item() - float(ref_loss)) < 2e-5
            _check_grads(tr, grads)
        cur = m.state_dict()
        for k, g in grads.items():
            err = (cur[k].cpu() - sd[k]).abs()
            if cfg["name"] != "sgd_plain" and it == 0:
                # Adam's first update is lr*sign(g): elements whose gradient is round-off noise may flip
                # (the update is g/(|g|+eps): a relative gradient error e moves it by about lr*e)
                noise = g.abs() <= 1e-3 * g.abs().max()
                err = torch.where(noise, torch.zeros_like(err), err)
            if it == 0:
                tol = 2e-5 if cfg["name"] == "sgd_plain" else 0.1 * cfg["lr"]
                assert int((err > tol).sum()) <= max(2, err.numel() // 100), (k, err.max().item())
        del before


def test_autograd_bridge_matches_trainer_and_torch_optim_steps():
    """The reference's loop shape: out = model(X); loss = criterion(out, y); loss.backward(); opt.step()."""
    dev = _dev()
    m = _lfan(MODS, dev, seed=2, p_drop=0.0)
    sd = synthetic.lfan_state_dict(2, MODS)
    X = synthetic.feature_windows(2, 300, seed=41, modalities=MODS)
    labels = torch.randint(0, 7, (2, 300, 1), generator=torch.Generator().manual_seed(42)).float()
    params = [p for p in m.parameters() if p.requires_grad]
    with torch.enable_grad():
        out = m({k: v.to(dev) for k, v in X.items()})
        assert out.requires_grad and out.shape == (2, 300, 7)
        loss = torch.nn.functional.cross_entropy(out.view(600, 7), labels.view(600).long().to(dev))
        loss.backward()
    ref_loss, grads, new_sd, _ = O.train_step(sd, X, labels, MODS, {"name": "sgd", "lr": 1e-2, "momentum": 0.9, "nesterov": True,
                                                                    "weight_decay": 1e-4}, None)
    assert abs(loss.item() - float(ref_loss)) < 2e-5
    named = dict(m.named_parameters())
    for k, g in grads.items():
        assert named[k].grad is not None, k
        _grad_close(named[k].grad.cpu(), g)
    opt = torch.optim.SGD(params, lr=1e-2, momentum=0.9, nesterov=True, weight_decay=1e-4)
    opt.step()
    cur = m.state_dict()
    for k in grads:
        d = (cur[k].cpu() - new_sd[k]).abs()
        assert int((d > 1e-5).sum()) <= max(2, d.numel() // 100) and d.max().item() < 1e-3, k
    # eval after training uses the updated weights (engines are re-packed)
    m.eval()
    with torch.no_grad():
        ev = m({k: v.to(dev) for k, v in X.items()}).cpu()
    want = O.lfan_forward({**new_sd}, X, MODS)
    assert (ev - want).abs().max().item() <= 2e-2


def test_ce_loss_and_optimizer_entry_points():
    from feature_vs_text_compound_emotion_b200 import _capi
    dev = _dev()
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(1000, 7, generator=g) * 3
    labels = torch.randint(0, 7, (1000,), generator=g)
    ref = torch.nn.functional.cross_entropy(logits, labels)
    with torch.enable_grad():
        lg = logits.clone().requires_grad_(True)
        torch.nn.functional.cross_entropy(lg, labels).backward()
    ld, yd = logits.to(dev), labels.to(dev)
    loss = torch.empty(1, device=dev)
    dl = torch.empty_like(ld)
    _capi.check(_capi.lib().cer_ce_loss(ld.data_ptr(), yd.data_ptr(), 1000, 7, loss.data_ptr(), dl.data_ptr(),
                                        _capi.current_stream_ptr()))
    assert abs(loss.item() - ref.item()) < 1e-5
    assert (dl.cpu() - lg.grad).abs().max().item() < 1e-7


def test_training_from_pixels_with_frozen_backbone_and_odd_widths():
    """(1) LFAN(video, vggish, bert).train(): frames go through the frozen IR-50 kernels (eval mode,
    see INTEGRATION.md), the head trains; compared with the oracle on the oracle's own embeddings.
    (2) a head with 39-/88-wide inputs (mfcc, egemaps) through the training GEMMs' unaligned paths."""
    dev = _dev()
    from feature_vs_text_compound_emotion_b200.models.model import LFAN
    from feature_vs_text_compound_emotion_b200.training import HeadTrainer
    mods = ["video", "vggish", "bert"]
    m = LFAN(backbone_settings=BS, output_dim=7, task="CLASSIFICATION", modality=mods, kernel_size=5, example_length=60,
             tcn_channel=synthetic.TCN_CHANNELS, modal_dim=32, num_heads=2, root_dir="", device=dev)
    m.init(visual_state_dict=synthetic.visual_backbone_state_dict(4))
    sd = synthetic.lfan_state_dict(4, mods)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).train()
    m.train_precision = "fp32"
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    f = synthetic.feature_windows(1, 60, seed=51, modalities=["vggish", "bert"])
    vid = synthetic.frames(60, seed=52).view(1, 60, 3, 40, 40)
    labels = torch.randint(0, 7, (1, 60, 1), generator=torch.Generator().manual_seed(53)).float()
    with torch.enable_grad():
        out = m({"video": vid.to(dev), "vggish": f["vggish"].to(dev), "bert": f["bert"].to(dev)})
        loss = torch.nn.functional.cross_entropy(out.view(60, 7), labels.view(60).long().to(dev))
        loss.backward()
    assert all(p.grad is None for p in m.spatial.parameters())
    # the head's gradients are checked on the embeddings the frozen kernels produced (with 60 rows the
    # batch-statistics BatchNorm amplifies the backbone's bf16 noise, which is a property of the model)
    with torch.no_grad():
        emb = m.spatial["visual"](vid.view(60, 3, 40, 40).to(dev)).cpu().view(1, 60, 512)
    assert F.cosine_similarity(emb[0], O.ir50_forward(sd, vid.view(60, 3, 40, 40), "spatial.visual.backbone."), dim=1).min() >= 0.999
    ref_loss, grads, _, _ = O.train_step(sd, {"video": emb, "vggish": f["vggish"], "bert": f["bert"]}, labels, mods,
                                         {"name": "sgd", "lr": 0.0}, None)
    assert abs(loss.item() - float(ref_loss)) < 2e-5
    named = dict(m.named_parameters())
    for k, g in grads.items():
        _grad_close(named[k].grad.cpu(), g)

    mods2 = ["mfcc", "egemaps"]
    m2 = _lfan(mods2, dev, seed=6, length=100, p_drop=0.0)
    tr = HeadTrainer(m2, 2, 100, precision="fp32")
    sd2 = synthetic.lfan_state_dict(6, mods2)
    X = synthetic.feature_windows(2, 100, seed=54, modalities=mods2)
    y = torch.randint(0, 7, (2, 100, 1), generator=torch.Generator().manual_seed(55)).float()
    logits = tr.forward({k: v.to(dev) for k, v in X.items()}, seed=1)
    loss2, dl = tr.cross_entropy(logits, y.to(dev))
    tr.backward(dl)
    ref2, grads2, _, _ = O.train_step(sd2, X, y, mods2, {"name": "sgd", "lr": 0.0}, None)
    assert abs(loss2.item() - float(ref2)) < 2e-5
    _check_grads(tr, grads2)


def test_eval_after_torch_optimizer_step_uses_new_weights():
    """eval -> train forward/backward + torch optimizer.step() -> eval: the second eval must see the
    updated weights and BatchNorm statistics (the engines hold packed copies of the old ones)."""
    dev = _dev()
    m = _lfan(MODS, dev, seed=7, p_drop=0.0)
    X = synthetic.feature_windows(2, 300, seed=71, modalities=MODS)
    labels = torch.randint(0, 7, (2, 300, 1), generator=torch.Generator().manual_seed(72))
    Xd = lambda: {k: v.to(dev) for k, v in X.items()}
    m.eval()
    with torch.no_grad():
        before = m(Xd()).clone()                 # builds the eval engines and the head CUDA graph
        assert torch.equal(m(Xd()), before)
    m.train()
    opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=0.5)
    with torch.enable_grad():
        loss = torch.nn.functional.cross_entropy(m(Xd()).view(600, 7), labels.view(600).to(dev))
        loss.backward()
    opt.step()
    m.eval()
    with torch.no_grad():
        after = m(Xd()).cpu()
    assert (after - before.cpu()).abs().max().item() > 1e-2          # not the stale engines
    want = O.lfan_forward({k: v.detach().cpu() for k, v in m.state_dict().items()}, {k: v.clone() for k, v in X.items()}, MODS)
    assert (after - want).abs().max().item() <= 2e-2
    # a second in-place update without any training forward in between is noticed as well
    with torch.no_grad():
        m.regressor.bias.add_(1.0)
        assert (m(Xd()).cpu() - after - 1.0).abs().max().item() < 1e-4


def test_ragged_batch_shares_buffers_and_stale_backward_raises():
    """The last partial batch of an epoch ((B', T) != (B, T)) adds a plan over the SAME flat parameter /
    gradient buffers; backward of an output whose saved activations were overwritten raises."""
    dev = _dev()
    m = _lfan(MODS, dev, seed=8, p_drop=0.0)
    sd = synthetic.lfan_state_dict(8, MODS)
    Xa = synthetic.feature_windows(3, 300, seed=81, modalities=MODS)
    Xb = synthetic.feature_windows(1, 300, seed=82, modalities=MODS)
    yb = torch.randint(0, 7, (1, 300, 1), generator=torch.Generator().manual_seed(83)).float()
    with torch.enable_grad():
        out_a = m({k: v.to(dev) for k, v in Xa.items()})
        tr = m.__dict__["_trainer"]
        flat = tr.params.data_ptr()
        out_b = m({k: v.to(dev) for k, v in Xb.items()})
        assert m.__dict__["_trainer"] is tr and tr.params.data_ptr() == flat and len(tr._plans) == 2
        with pytest.raises(RuntimeError, match="before backward"):
            out_a.sum().backward()
        loss = torch.nn.functional.cross_entropy(out_b.view(300, 7), yb.view(300).long().to(dev))
        loss.backward()
    ref_loss, grads, _, _ = O.train_step(sd, Xb, yb, MODS, {"name": "sgd", "lr": 0.0}, None)
    assert abs(loss.item() - float(ref_loss)) < 2e-5
    named = dict(m.named_parameters())
    for k, g in grads.items():
        _grad_close(named[k].grad.cpu(), g)


def test_ce_loss_ignore_index_and_out_of_range_labels():
    from feature_vs_text_compound_emotion_b200 import _capi
    dev = _dev()
    g = torch.Generator().manual_seed(6)
    logits = torch.randn(777, 7, generator=g) * 2
    labels = torch.randint(0, 7, (777,), generator=g)
    labels[::5] = -100
    ref_in = logits.clone().requires_grad_(True)
    with torch.enable_grad():
        ref = torch.nn.functional.cross_entropy(ref_in, labels)
        ref.backward()
    ld = logits.to(dev)                                    # device operands stay referenced until the results are read

    def ce(lab, want_grad):
        lab_d = lab.to(dev)
        loss = torch.empty(1, device=dev)
        dl = torch.full((777, 7), 9.0, device=dev) if want_grad else None
        _capi.check(_capi.lib().cer_ce_loss(ld.data_ptr(), lab_d.data_ptr(), 777, 7, loss.data_ptr(), None if dl is None else dl.data_ptr(),
                                            _capi.current_stream_ptr()))
        torch.cuda.synchronize()
        return loss.cpu(), None if dl is None else dl.cpu()

    loss, dl = ce(labels, True)
    assert abs(loss.item() - ref.item()) < 1e-5
    assert (dl - ref_in.grad).abs().max().item() < 1e-7
    # out-of-range labels are never used as an index: they behave like ignored rows
    lab2 = labels.clone()
    lab2[labels == -100] = 7
    lab2[0] = -3 if labels[0] == -100 else lab2[0]
    assert abs(ce(lab2, False)[0].item() - ref.item()) < 1e-5
    assert torch.isnan(ce(torch.full((777,), -100), False)[0]).item()


# ----------------------------------------------------------------------------------------------
# TF32 tensor-core mode (HeadTrainer's default: the TCN convolutions' GEMMs see TF32-rounded operands).
# What TF32 does to THIS model, measured with the oracle on the CPU (no GPU involved):
#   * O.train_step(..., tf32=True) vs its own fp32 run: gradients differ by 2.5 % (median over the tensors)
#     to 7 % (worst) in relative L2, logits by 2e-3 -- BatchNorm1d with batch statistics and LayerNorm
#     amplify the 5e-4 operand rounding ~30x;
#   * two TF32 runs whose weights differ by fp32 round-off (2e-7 relative) differ by 1e-3 in the logits:
#     values sitting on a TF32 rounding boundary flip, and the flips are amplified the same way.
# So no TF32 implementation can be pinned to another tighter than that.  The bars below are those
# measured deviations with margin, plus direction (cosine) and convergence checks; the exact-fp32 mode
# above carries the tight parity.
# ----------------------------------------------------------------------------------------------
_TF32_STATS = []


def _tf32_grads_close(mine: dict, ref: dict, what: str, norms=None):
    rels = []
    for k, g in ref.items():
        a, b = mine[k].double().flatten(), g.double().flatten()
        norm = float(b.norm()) if norms is None else norms[k]
        rel = float((a - b).norm()) / max(norm, 1e-30)
        cos = float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))
        _TF32_STATS.append((what + ": " + k, rel, cos))
        assert rel <= 0.25, (what, k, rel)
        assert cos >= 0.97, (what, k, cos)
        rels.append(rel)
    rels.sort()
    assert rels[len(rels) // 2] <= 0.06, (what, "median", rels[len(rels) // 2])
    return rels


@pytest.fixture(scope="module", autouse=True)
def _dump_tf32_stats():
    yield
    path = os.environ.get("CER_GRAD_STATS")
    if path and _TF32_STATS:
        import json
        with open(path.replace(".json", "_tf32.json"), "w") as f:
            json.dump(_TF32_STATS, f)


def test_tf32_step_vs_tf32_oracle_and_fp32_mode():
    """TF32 mode, dropout ON, against the oracle run with TF32-rounded conv operands (same masks, same
    cvt.rna rounding) and against this repo's exact fp32 mode on the same inputs: logits within 5e-3, loss
    within 2e-3, every gradient within the TF32 deviation of this model (see the block comment above)."""
    from feature_vs_text_compound_emotion_b200.training import HeadTrainer
    dev = _dev()
    sd = synthetic.lfan_state_dict(3, MODS)
    X = synthetic.feature_windows(2, 300, seed=21, modalities=MODS)
    labels = torch.randint(0, 7, (2, 300, 1), generator=torch.Generator().manual_seed(22)).float()
    seed = 0xC0FFEE
    outs = {}
    for prec in ("tf32", "fp32"):
        m = _lfan(MODS, dev, seed=3, precision=prec)
        tr = HeadTrainer(m, 2, 300, precision=prec)
        logits = tr.forward({k: v.to(dev) for k, v in X.items()}, seed=seed)
        loss, dl = tr.cross_entropy(logits, labels.to(dev))
        tr.backward(dl)
        outs[prec] = (logits.cpu(), loss.item(), {k: tr.grad(k).cpu().clone() for k in tr.names})
    lt, losst, gt = outs["tf32"]
    lf, lossf, gf = outs["fp32"]
    P = {k: sd[k] for k in O.trainable_names(sd)}
    buffers = {k: v.clone() for k, v in sd.items() if k.startswith("bn.") and k not in P}
    want = O.head_forward_train(P, buffers, {k: v.squeeze(1) for k, v in X.items()}, MODS, seed=seed, tf32=True)
    assert (lt - want).abs().max().item() < 5e-3
    ref_loss, grads, _, _ = O.train_step(sd, X, labels, MODS, {"name": "sgd", "lr": 0.0}, None, seed=seed, tf32=True)
    assert abs(losst - float(ref_loss)) < 2e-3
    _tf32_grads_close(gt, grads, "vs tf32 oracle")
    assert (lt - lf).abs().max().item() < 5e-3 and not torch.equal(lt, lf)
    assert abs(losst - lossf) < 2e-3
    _tf32_grads_close(gt, {k: gf[k] for k in grads}, "vs fp32 mode")


def test_tf32_two_sgd_steps_vs_reference_golden(golden_dir):
    """The default (TF32) training step against the REFERENCE's two fp32 SGD steps: loss within 2e-3, gradient
    norms within 5 %, gradients within the TF32 deviation, BatchNorm running statistics within 1e-3."""
    from feature_vs_text_compound_emotion_b200.training import HeadTrainer
    dev = _dev()
    g = torch.load(os.path.join(golden_dir, "train_b2.pt"))
    mods = g["modalities"]
    m = _lfan(mods, dev, g["weights_seed"], p_drop=0.0, precision="tf32")
    tr = HeadTrainer(m, 2, 300, optimizer=g["opt"])
    assert tr.precision == "tf32"
    X = {k: v.to(dev) for k, v in synthetic.feature_windows(2, 300, seed=g["x_seed"], modalities=mods).items()}
    labels = torch.randint(0, 7, (2, 300, 1), generator=torch.Generator().manual_seed(g["label_seed"])).float().to(dev)
    for i, step in enumerate(g["steps"]):
        loss = tr.step(X, labels)
        assert abs(loss.item() - step["loss"]) < 2e-3
        if i == 0:       # same parameters as the reference: tensor-by-tensor (from step 2 on the parameters themselves differ)
            small = {k: tr.grad(k).cpu() for k in step["grad_small"]}
            _tf32_grads_close(small, step["grad_small"], "vs reference golden", norms=step["grad_norm"])
            sample = {k: tr.grad(k).cpu().flatten()[::997] for k in step["grad_sample"]}
            _tf32_grads_close(sample, step["grad_sample"], "vs reference golden (1/997 sample)")
        for k, gn in step["grad_norm"].items():
            assert abs(float(tr.grad(k).double().norm()) - gn) <= (5e-2 if i == 0 else 0.15) * gn + 1e-7, k
        sd = m.state_dict()
        for k, v in step["bn"].items():
            assert (sd[k].cpu().float() - v.float()).abs().max().item() < 1e-3, k


def test_tf32_and_fp32_training_converge_alike():
    """30 AdamW steps on the same batches in both precisions: the loss curves stay together (TF32 noise
    does not change the optimisation), and TF32 really is the tensor-core path (different bits)."""
    from feature_vs_text_compound_emotion_b200.training import HeadTrainer
    dev = _dev()
    X = [synthetic.feature_windows(4, 300, seed=91 + i, modalities=MODS) for i in range(2)]
    y = [torch.randint(0, 7, (4, 300, 1), generator=torch.Generator().manual_seed(95 + i)).float() for i in range(2)]
    curves = {}
    for prec in ("tf32", "fp32"):
        m = _lfan(MODS, dev, seed=9, precision=prec)
        tr = HeadTrainer(m, 4, 300, optimizer={"name": "adamw", "lr": 1e-3, "weight_decay": 1e-4}, seed=5, precision=prec)
        curves[prec] = [tr.step({k: v.to(dev) for k, v in X[i % 2].items()}, y[i % 2].to(dev)).item() for i in range(30)]
    a, b = curves["tf32"], curves["fp32"]
    assert a[0] != b[0] and abs(a[0] - b[0]) < 2e-3
    assert b[-1] < 0.8 * b[0]                                   # it does learn these batches
    assert abs(a[-1] - b[-1]) <= 0.05 * b[-1], (a[-1], b[-1])
