"""Generate tests/golden/*.pt by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):
    python oracle/gen_golden.py

What it pins (SURVEY.md section 8c -- the reference has no golden vectors of its own, so the
reference modules, imported as they are, are the source of truth):
  * ir50_n4.pt      -- VisualBackbone forward on 4 synthetic frames (weights: synthetic seed 0)
  * head_b2.pt      -- LFAN(cnn_res50,vggish,bert) forward, B=2 x T=300  (BASELINE cfg 1 shape)
  * lfan_b1.pt      -- LFAN(video,vggish,bert) forward from pixels, B=1 x T=300
  * state_keys.json -- the reference's own state_dict key/shape listing (534 keys)
  * vggish_n6.pt    -- VGGish forward on 6 synthetic log-mel examples (+ its 18 state_dict keys)
  * lfan_logmel_b1.pt -- LFAN(video,logmel,bert) forward from pixels and log-mel, B=1 x T=40
  * train_b2.pt     -- two SGD-nesterov steps of the head in train mode (Dropout p=0): loss, gradients, BN stats
  * eval_transform.pt -- the reference's eval transform classes (PIL resize 48, crop 40, normalise) on seeded uint8 frames
  * logmel.pt       -- mel_features.log_mel_spectrogram + my_frame (the reference code) on a seeded 3 s waveform
  * heads.pt        -- CAN / JMT / MT forward from pixels (B=2 x T=24) and their state_dict key listings
  * heads_t300.pt   -- the same heads at the reference's window length (B=2 x T=300)
  * heads_train.pt  -- one training step (loss, gradients, BatchNorm statistics) of CAN / JMT / MT on pre-encoded features
  * attention_maps.pt -- MultimodalTransformerEncoder.get_attention_maps on seeded encoder outputs
  * windowing.json  -- Trainer.windowing outputs for a set of lengths
Weights are NOT stored: they are regenerated from the seed by
feature_vs_text_compound_emotion_b200.synthetic (identical on every machine), and are loaded
into the reference modules with strict=True here, which also proves the key layout.
"""
import json
import os
import sys
import tempfile
import warnings

import numpy as np
import torch
import torch._dynamo  # noqa: F401  (torch.optim pulls it in; must be imported before the reference edits sys.path)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from feature_vs_text_compound_emotion_b200 import synthetic  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _heads(TH, ref_configs, tmp, with_keys=True):
    """CAN / JMT / MT from pixels, B = 2 windows of TH frames, through the reference modules."""
    from models.model import CAN, JMT
    ts = ref_configs.config["tcn_settings"]
    assert {k: ts[k] for k in synthetic.TCN_SETTINGS} == synthetic.TCN_SETTINGS
    heads = {}
    cmods = ["video", "vggish", "bert"]
    can = CAN(task="CLASSIFICATION", modalities=cmods, tcn_settings=ts, backbone_settings=ref_configs.config["backbone_settings"],
              output_dim=7, root_dir=tmp, device="cpu")
    csd = synthetic.can_state_dict(0, cmods)
    assert list(can.state_dict()) == list(csd)
    can.load_state_dict(csd, strict=True)
    can.eval()
    fh = synthetic.feature_windows(2, TH, seed=705, modalities=["vggish", "bert"])
    Xc = {"video": synthetic.frames(2 * TH, seed=706).view(2, TH, 3, 40, 40), "vggish": fh["vggish"], "bert": fh["bert"]}
    heads["CAN"] = {"modalities": cmods, "feat_seed": 705, "frame_seed": 706, "T": TH, "out": can({k: v.clone() for k, v in Xc.items()})}
    if with_keys:
        heads["CAN"]["keys"] = {k: list(v.shape) for k, v in can.state_dict().items()}
    for name in ("JMT", "MT"):
        jmods = ["video", "vggish"]
        jm = JMT(task="CLASSIFICATION", modalities=jmods, tcn_settings=ts, backbone_settings=ref_configs.config["backbone_settings"],
                 output_dim=7, root_dir=tmp, device="cpu", model_name=name)
        jsd = synthetic.jmt_state_dict(0, jmods, model_name=name)
        assert list(jm.state_dict()) == list(jsd)
        jm.load_state_dict(jsd, strict=True)
        jm.eval()
        Xj = {"video": synthetic.frames(2 * TH, seed=707).view(2, TH, 3, 40, 40),
              "vggish": synthetic.feature_windows(2, TH, seed=708, modalities=["vggish"])["vggish"]}
        heads[name] = {"modalities": jmods, "feat_seed": 708, "frame_seed": 707, "T": TH, "out": jm({k: v.clone() for k, v in Xj.items()})}
        if with_keys:
            heads[name]["keys"] = {k: list(v.shape) for k, v in jm.state_dict().items()}
    return heads


def _heads_train(ref_configs, tmp):
    """One training step (model.train(), mean cross-entropy, backward) of the REFERENCE CAN / JMT / MT on pre-encoded
    features, B = 2 x T = 40.  The frozen visual backbone is replaced by a Flatten stub fed with [B, T, 512, 1, 1]
    "frames", so that the head sees exactly the 512-d embeddings this repo's training path is given (the reference
    would otherwise also flip IR-50's BatchNorms to batch statistics); Dropout p = 0 (torch's Philox masks cannot be
    reproduced -- dropout ON is checked against the oracle, which restates the kernels' mask hash)."""
    import torch.nn.functional as F
    from models.model import CAN, JMT
    ts = ref_configs.config["tcn_settings"]
    out = {}
    for name in ("CAN", "JMT", "MT"):
        mods = ["video", "vggish", "bert"] if name == "CAN" else ["video", "vggish"]
        if name == "CAN":
            m = CAN(task="CLASSIFICATION", modalities=mods, tcn_settings=ts, backbone_settings=ref_configs.config["backbone_settings"],
                    output_dim=7, root_dir=tmp, device="cpu")
            sd = synthetic.can_state_dict(0, mods)
        else:
            m = JMT(task="CLASSIFICATION", modalities=mods, tcn_settings=ts, backbone_settings=ref_configs.config["backbone_settings"],
                    output_dim=7, root_dir=tmp, device="cpu", model_name=name)
            sd = synthetic.jmt_state_dict(0, mods, model_name=name)
        m.load_state_dict(sd, strict=True)
        m.spatial["visual"] = torch.nn.Flatten()
        m.train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        B, T = 2, 40
        g = torch.Generator().manual_seed(811)
        dims = {"video": 512, "vggish": 128, "bert": 768}
        feats = {k: torch.randn(B, T, dims[k], generator=g) for k in mods}
        labels = torch.randint(0, 7, (B, T, 1), generator=g)
        X = {k: (v.view(B, T, 512, 1, 1) if k == "video" else v.unsqueeze(1)).clone() for k, v in feats.items()}
        with torch.enable_grad():
            logits = m(X)
            loss = F.cross_entropy(logits.reshape(-1, 7), labels.reshape(-1))
            loss.backward()
        rec = {"modalities": mods, "seed": 811, "B": B, "T": T, "loss": float(loss), "logits": logits.detach().clone(),
               "grad_norm": {}, "grad_small": {}, "grad_sample": {}, "none_grad": [], "bn": {}}
        for k, p in m.named_parameters():
            if k.startswith("spatial.") or ".net." in k:
                continue
            if p.grad is None:
                rec["none_grad"].append(k)
                continue
            rec["grad_norm"][k] = float(p.grad.double().norm())
            if p.grad.numel() <= 4096:
                rec["grad_small"][k] = p.grad.detach().clone()
            else:
                rec["grad_sample"][k] = p.grad.detach().flatten()[::97].clone()
        for k, v in m.state_dict().items():
            if (k.startswith("bn.") or k.startswith("bn1.")) and ("running" in k):
                rec["bn"][k] = v.detach().clone()
        out[name] = rec
        print("heads_train", name, rec["loss"], len(rec["grad_norm"]), "grads,", len(rec["none_grad"]), "without grad")
    return out


def only_heads_train():
    """python oracle/gen_golden.py headstrain -- just tests/golden/heads_train.pt."""
    torch.manual_seed(0)
    import configs as ref_configs
    tmp = tempfile.mkdtemp()
    torch.save(synthetic.visual_backbone_state_dict(seed=0), os.path.join(tmp, "res50_ir_0.887.pth"))
    torch.save(_heads_train(ref_configs, tmp), os.path.join(OUT, "heads_train.pt"))


def only_heads_t300():
    """python oracle/gen_golden.py heads300 -- just tests/golden/heads_t300.pt (the other fixtures are untouched)."""
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    import configs as ref_configs
    tmp = tempfile.mkdtemp()
    torch.save(synthetic.visual_backbone_state_dict(seed=0), os.path.join(tmp, "res50_ir_0.887.pth"))
    heads = _heads(300, ref_configs, tmp, with_keys=False)
    torch.save(heads, os.path.join(OUT, "heads_t300.pt"))
    print("heads_t300", {k: tuple(v["out"].shape) for k, v in heads.items()})


def main():
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    import configs as ref_configs
    from models.backbone import VisualBackbone
    from models.model import LFAN

    os.makedirs(OUT, exist_ok=True)

    # ---- IR-50 -----------------------------------------------------------------------
    vsd = synthetic.visual_backbone_state_dict(seed=0)
    vb = VisualBackbone(use_pretrained=False)
    vb.load_state_dict(vsd, strict=True)
    vb.eval()
    x = synthetic.frames(4, seed=1234)
    emb = vb(x)
    # a few intermediate activations, to localise a failure to a stage
    h = vb.backbone.input_layer(x)
    taps = {"stem": h[:, :, ::13, ::13].clone()}
    for i, unit in enumerate(vb.backbone.body):
        h = unit(h)
        if i in (2, 3, 6, 7, 20, 21, 23):
            taps[f"body{i}"] = h[:, ::17, ::3, ::3].clone()
    torch.save({"x_seed": 1234, "n": 4, "weights_seed": 0, "emb": emb, "taps": taps},
               os.path.join(OUT, "ir50_n4.pt"))
    print("ir50", emb.shape, float(emb.abs().mean()), {k: float(v.abs().mean()) for k, v in taps.items()})

    # ---- head only (cfg 1 shape) ---------------------------------------------------------
    mods = ["cnn_res50", "vggish", "bert"]
    tmp = tempfile.mkdtemp()
    torch.save(vsd, os.path.join(tmp, "res50_ir_0.887.pth"))
    model = LFAN(backbone_settings=ref_configs.config["backbone_settings"], output_dim=7,
                 task="CLASSIFICATION", modality=mods, kernel_size=5, example_length=300,
                 tcn_channel=ref_configs.config["tcn"]["channels"], modal_dim=32, num_heads=2,
                 root_dir=tmp, device="cpu")
    model.init()
    hsd = synthetic.lfan_state_dict(seed=0, modalities=mods)
    model.load_state_dict(hsd, strict=True)
    model.eval()
    X = synthetic.feature_windows(2, 300, seed=1234, modalities=mods)
    logits = model({k: v.clone() for k, v in X.items()})
    torch.save({"x_seed": 1234, "batch": 2, "weights_seed": 0, "modalities": mods, "logits": logits},
               os.path.join(OUT, "head_b2.pt"))
    print("head", logits.shape, float(logits.abs().mean()))

    # ---- full LFAN from pixels -------------------------------------------------------------
    mods3 = ["video", "vggish", "bert"]
    full = LFAN(backbone_settings=ref_configs.config["backbone_settings"], output_dim=7,
                task="CLASSIFICATION", modality=mods3, kernel_size=5, example_length=300,
                tcn_channel=ref_configs.config["tcn"]["channels"], modal_dim=32, num_heads=2,
                root_dir=tmp, device="cpu")
    full.init()
    fsd = synthetic.lfan_state_dict(seed=0, modalities=mods3)
    full.load_state_dict(fsd, strict=True)
    full.eval()
    keys = {k: list(v.shape) for k, v in full.state_dict().items()}
    assert list(keys) == list(fsd), "synthetic key order differs from the reference's"
    json.dump(keys, open(os.path.join(OUT, "state_keys.json"), "w"), indent=0)
    feats = synthetic.feature_windows(1, 300, seed=77, modalities=["vggish", "bert"])
    vid = synthetic.frames(300, seed=78).view(1, 300, 3, 40, 40)
    Xf = {"video": vid, "vggish": feats["vggish"], "bert": feats["bert"]}
    lg = full({k: v.clone() for k, v in Xf.items()})
    torch.save({"feat_seed": 77, "frame_seed": 78, "weights_seed": 0, "modalities": mods3, "logits": lg},
               os.path.join(OUT, "lfan_b1.pt"))
    print("lfan", lg.shape, float(lg.abs().mean()), len(keys), "keys")

    # ---- VGGish + LFAN with the inline `logmel` modality ---------------------------------------
    from models.backbone import VGGish
    asd = synthetic.vggish_state_dict(seed=0)
    vg = VGGish()
    vg.load_state_dict(asd, strict=True)
    vg.eval()
    xa = synthetic.logmel_patches(6, seed=1234)
    ya = vg(xa)
    torch.save({"x_seed": 1234, "n": 6, "weights_seed": 0, "emb": ya, "keys": {k: list(v.shape) for k, v in vg.state_dict().items()}},
               os.path.join(OUT, "vggish_n6.pt"))
    print("vggish", ya.shape, float(ya.abs().mean()))

    modsl = ["video", "logmel", "bert"]
    TL = 40
    torch.save(asd, os.path.join(tmp, "vggish.pth"))
    lm = LFAN(backbone_settings=ref_configs.config["backbone_settings"], output_dim=7,
              task="CLASSIFICATION", modality=modsl, kernel_size=5, example_length=TL,
              tcn_channel=synthetic.TCN_CHANNELS, modal_dim=32, num_heads=2, root_dir=tmp, device="cpu")
    lm.init()
    lsd = synthetic.lfan_state_dict(seed=0, modalities=modsl)
    assert list(lm.state_dict()) == list(lsd), "synthetic key order differs from the reference's (logmel)"
    lm.load_state_dict(lsd, strict=True)
    lm.eval()
    fl = synthetic.feature_windows(1, TL, seed=91, modalities=["bert"])
    Xl = {"video": synthetic.frames(TL, seed=92).view(1, TL, 3, 40, 40),
          "logmel": synthetic.logmel_patches(TL, seed=93).view(1, TL, 96, 64).permute(0, 3, 1, 2).contiguous(),
          "bert": fl["bert"]}
    ll = lm({k: v.clone() for k, v in Xl.items()})
    torch.save({"length": TL, "bert_seed": 91, "frame_seed": 92, "logmel_seed": 93, "weights_seed": 0,
                "modalities": modsl, "logits": ll, "keys": {k: list(v.shape) for k, v in lm.state_dict().items()}},
               os.path.join(OUT, "lfan_logmel_b1.pt"))
    print("lfan logmel", ll.shape, float(ll.abs().mean()), len(lsd), "keys")

    # ---- fusion-head training step (cfg 4), dropout p = 0 so that it is reproducible ---------
    sys.path[:] = [q for q in sys.path if isinstance(q, str)]
    model.load_state_dict(hsd, strict=True)
    model.train()
    for mod in model.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    Xt = synthetic.feature_windows(2, 300, seed=4321, modalities=mods)
    labels = torch.randint(0, 7, (2, 300, 1), generator=torch.Generator().manual_seed(4322)).float()
    opt_cfg = {"name": "sgd", "lr": 1e-2, "momentum": 0.9, "dampening": 0.0, "weight_decay": 1e-4, "nesterov": True}
    params = [q for q in model.parameters() if q.requires_grad]
    opt = torch.optim.SGD(params, lr=opt_cfg["lr"], momentum=0.9, dampening=0.0, weight_decay=1e-4, nesterov=True)
    rec = {"x_seed": 4321, "label_seed": 4322, "weights_seed": 0, "modalities": mods, "opt": opt_cfg, "steps": []}
    with torch.enable_grad():
        for it in range(2):
            opt.zero_grad(set_to_none=True)
            outputs = model({k: v.clone() for k, v in Xt.items()})
            loss = torch.nn.functional.cross_entropy(outputs.contiguous().view(600, 7), labels.view(600).long())
            loss.backward()
            grads = {k: q.grad.detach().clone() for k, q in model.named_parameters() if q.grad is not None}
            opt.step()
            after = model.state_dict()
            step = {"loss": float(loss), "grad_sum": {k: float(g.double().sum()) for k, g in grads.items()},
                    "grad_norm": {k: float(g.double().norm()) for k, g in grads.items()},
                    "grad_small": {k: g for k, g in grads.items() if g.numel() <= 4096},
                    "grad_sample": {k: g.flatten()[::997].clone() for k, g in grads.items() if g.numel() > 4096},
                    "param_sample": {k: v.detach().flatten()[::997].clone() for k, v in after.items() if v.dtype.is_floating_point},
                    "bn": {k: v.clone() for k, v in after.items() if k.startswith("bn.")}}
            rec["steps"].append(step)
            print("train step", it, float(loss), len(grads), "grads")
    torch.save(rec, os.path.join(OUT, "train_b2.pt"))
    torch.set_grad_enabled(False)

    # ---- eval input transform (base/dataset.py:503-510) through the reference's own classes ----
    import torchvision.transforms as TT
    from base.transforms3D import (GroupCenterCrop, GroupNormalize, GroupNumpyToPILImage, GroupScale, Stack,
                                   ToTorchFormatTensor)
    tr = TT.Compose([GroupNumpyToPILImage(use_inverse=False), GroupScale(48), GroupCenterCrop(40), Stack(),
                     ToTorchFormatTensor(), GroupNormalize([0.5, 0.5, 0.5], [0.5, 0.5, 0.5])])
    cases = {}
    for name, (n, hh, ww, sd_) in {"256": (3, 256, 256, 501), "112": (2, 112, 112, 502), "64x80": (2, 64, 80, 503),
                                   "40": (1, 40, 40, 504)}.items():
        raw = synthetic.raw_frames_u8(n, seed=sd_, h=hh, w=ww)
        cases[name] = {"n": n, "h": hh, "w": ww, "seed": sd_, "out": tr(raw.numpy())}
    import PIL
    torch.save({"cases": cases, "pillow": PIL.__version__}, os.path.join(OUT, "eval_transform.pt"))
    print("eval transform", {k: tuple(v["out"].shape) for k, v in cases.items()}, "Pillow", PIL.__version__)

    # ---- log-mel front end (abaw5_pre_processing/base/vggish/mel_features.py), reference code itself ----
    sys.path.insert(0, os.path.join(REF, "abaw5_pre_processing"))
    from base.vggish import mel_features as MF
    wave = synthetic.waveform(3.0, seed=601)
    lm = MF.log_mel_spectrogram(wave.double().numpy(), audio_sample_rate=16000, log_offset=0.01, window_length_secs=0.025,
                                hop_length_secs=0.010, num_mel_bins=64, lower_edge_hertz=125, upper_edge_hertz=7500)
    hop_sec = 1.0 / 30.0                                   # one example per video frame at 30 fps (audio.py:126-127)
    ex = MF.my_frame(lm, window_length=96, hop_length=hop_sec * 100.0)
    torch.save({"seconds": 3.0, "seed": 601, "hop_sec": hop_sec, "log_mel": torch.from_numpy(lm).float(),
                "n_examples": int(ex.shape[0]), "example_sum": torch.from_numpy(ex.sum(axis=(1, 2))),
                "example_7": torch.from_numpy(ex[7]).float(), "example_last": torch.from_numpy(ex[-1]).float()},
               os.path.join(OUT, "logmel.pt"))
    print("logmel", lm.shape, ex.shape)

    # ---- alternative heads CAN / JMT / MT (models/model.py:529-684, :895-1167) ------------------
    heads = _heads(24, ref_configs, tmp)
    torch.save(heads, os.path.join(OUT, "heads.pt"))
    print("heads", {k: (tuple(v["out"].shape), len(v["keys"])) for k, v in heads.items()})
    heads = _heads(300, ref_configs, tmp, with_keys=False)         # the reference's window length
    torch.save(heads, os.path.join(OUT, "heads_t300.pt"))
    with torch.enable_grad():
        torch.save(_heads_train(ref_configs, tmp), os.path.join(OUT, "heads_train.pt"))
    print("heads_t300", {k: tuple(v["out"].shape) for k, v in heads.items()})

    # ---- attention maps of the cross-modal encoder (transformer.py:211-215) ---------------------
    from models.transformer import MultimodalTransformerEncoder
    enc = MultimodalTransformerEncoder(modalities=mods3, input_dim={"video": 128, "vggish": 32, "bert": 128}, modal_dim=32,
                                       num_heads=2, dropout=0.1)
    enc.load_state_dict({k[len("fusion."):]: v for k, v in fsd.items() if k.startswith("fusion.")}, strict=True)
    enc.eval()
    xa_ = {m: torch.randn(2, 50, d, generator=torch.Generator().manual_seed(3)) for m, d in zip(mods3, (128, 32, 128))}
    torch.save({"seed": 3, "T": 50, "maps": enc.get_attention_maps(xa_)[0]}, os.path.join(OUT, "attention_maps.pt"))

    # ---- windowing (trainer.py imports pynvml/munch, absent here: exec the one function) ----
    src = open(os.path.join(REF, "trainer.py")).read()
    start = src.index("    def windowing(x, window_length, hop_length)")
    body = src[start:]
    ns = {"np": np, "List": list}
    exec("import numpy as np\nfrom typing import List\n" + "\n".join(l[4:] for l in body.splitlines()), ns)
    cases = {}
    for L in (1, 150, 299, 300, 301, 499, 500, 501, 700, 701, 1234, 3000):
        cases[str(L)] = [[int(w[0]), int(w[-1]), len(w)] for w in ns["windowing"](np.arange(L), 300, 200)]
    json.dump(cases, open(os.path.join(OUT, "windowing.json"), "w"))
    print("windowing", {k: len(v) for k, v in cases.items()})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "heads300":
        only_heads_t300()
    elif len(sys.argv) > 1 and sys.argv[1] == "headstrain":
        only_heads_train()
    else:
        main()
