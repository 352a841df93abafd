"""End-to-end GPU parity of the drop-in modules against the reference-generated goldens.
Tolerances are BASELINE.json's: embedding cosine >= 0.999, logit max-abs <= 2e-2, argmax
agreement >= 99.5 %."""
import os
import warnings

import pytest
import torch
import torch.nn.functional as F

from feature_vs_text_compound_emotion_b200 import synthetic
from oracle import lfan_oracle as O

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
warnings.filterwarnings("ignore")
BS = {"visual_state_dict": "res50_ir_0.887", "audio_state_dict": "vggish"}


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _lfan(mods, dev, seed=0, length=300):
    from feature_vs_text_compound_emotion_b200.models.model import LFAN
    m = LFAN(backbone_settings=BS, output_dim=7, task="CLASSIFICATION", modality=mods, kernel_size=5,
             example_length=length, tcn_channel=synthetic.TCN_CHANNELS, modal_dim=32, num_heads=2, root_dir="",
             device=dev)
    m.init(visual_state_dict=synthetic.visual_backbone_state_dict(seed) if "video" in mods else None,
           audio_state_dict=synthetic.vggish_state_dict(seed) if "logmel" in mods else None)
    m.load_state_dict(synthetic.lfan_state_dict(seed, mods), strict=True)
    return m.to(dev).eval()


def test_head_only_vs_golden(golden_dir):
    """BASELINE config 1 shape: B=2 x T=300 pre-extracted features."""
    dev = _dev()
    g = torch.load(os.path.join(golden_dir, "head_b2.pt"))
    mods = g["modalities"]
    m = _lfan(mods, dev, g["weights_seed"])
    X = {k: v.to(dev) for k, v in synthetic.feature_windows(g["batch"], 300, seed=g["x_seed"], modalities=mods).items()}
    out = m(X).cpu()
    assert out.shape == (2, 300, 7)
    assert (out - g["logits"]).abs().max().item() <= 2e-2      # BASELINE.json tolerance
    assert (out - g["logits"]).abs().max().item() <= 5e-3      # TF32 head: ~1e-3 in practice
    assert (out.argmax(-1) == g["logits"].argmax(-1)).float().mean().item() >= 0.995


def test_full_lfan_from_pixels_vs_golden(golden_dir):
    dev = _dev()
    g = torch.load(os.path.join(golden_dir, "lfan_b1.pt"))
    mods = g["modalities"]
    m = _lfan(mods, dev, g["weights_seed"])
    feats = synthetic.feature_windows(1, 300, seed=g["feat_seed"], modalities=["vggish", "bert"])
    X = {"video": synthetic.frames(300, seed=g["frame_seed"]).view(1, 300, 3, 40, 40).to(dev),
         "vggish": feats["vggish"].to(dev), "bert": feats["bert"].to(dev)}
    out = m(X).cpu()
    err = (out - g["logits"]).abs().max().item()
    agree = (out.argmax(-1) == g["logits"].argmax(-1)).float().mean().item()
    assert err <= 2e-2, err
    assert agree >= 0.995, agree


def test_visual_backbone_module_cosine():
    dev = _dev()
    from feature_vs_text_compound_emotion_b200.models.backbone import VisualBackbone
    sd = synthetic.visual_backbone_state_dict(3)
    vb = VisualBackbone(use_pretrained=False)
    vb.load_state_dict(sd, strict=True)
    vb = vb.to(dev).eval()
    x = synthetic.frames(24, seed=8)
    emb = vb(x.to(dev)).cpu()
    ref = O.ir50_forward(sd, x, "backbone.")
    assert F.cosine_similarity(emb, ref, dim=1).min().item() >= 0.999
    assert torch.equal(vb.extract(x.to(dev)).cpu(), emb)


def test_submodules_keep_reference_layouts():
    """TemporalConvNet takes [B,C,T]; MultimodalTransformerEncoder takes a dict of [B,T,D]."""
    dev = _dev()
    mods = ["cnn_res50", "vggish", "bert"]
    m = _lfan(mods, dev)
    sd = synthetic.lfan_state_dict(0, mods)
    x = torch.randn(2, 128, 300, generator=torch.Generator().manual_seed(1))
    got = m.temporal["vggish"](x.to(dev)).cpu()
    want = O.tcn_forward(sd, "temporal.vggish.", x)
    assert got.shape == (2, 32, 300) and (got - want).abs().max().item() < 5e-3
    enc = {k: torch.randn(2, 300, d, generator=torch.Generator().manual_seed(2)) for k, d in
           zip(mods, (128, 32, 128))}
    fused = m.fusion({k: v.to(dev) for k, v in enc.items()}).cpu()
    assert (fused - O.fusion_forward(sd, "fusion.", enc, mods)).abs().max().item() < 1e-4


def test_argmax_agreement_over_many_frames():
    """600 frames from pixels, different weights seed: the 99.5 % argmax bar with margin."""
    dev = _dev()
    mods = ["video", "vggish", "bert"]
    m = _lfan(mods, dev, seed=5)
    sd = synthetic.lfan_state_dict(5, mods)
    feats = synthetic.feature_windows(2, 300, seed=41, modalities=["vggish", "bert"])
    vid = synthetic.frames(600, seed=42).view(2, 300, 3, 40, 40)
    X = {"video": vid, "vggish": feats["vggish"], "bert": feats["bert"]}
    ref = O.lfan_forward(sd, {k: v.clone() for k, v in X.items()}, mods)
    out = m({k: v.to(dev) for k, v in X.items()}).cpu()
    assert (out - ref).abs().max().item() <= 2e-2
    assert (out.argmax(-1) == ref.argmax(-1)).float().mean().item() >= 0.995


def test_long_video_windowing_dedup_and_stitch():
    """501-frame video: unique frames encoded once, 3 windows batched through the head, stitched on
    device -- equals the reference procedure (every window from scratch, overlap-averaged)."""
    dev = _dev()
    from feature_vs_text_compound_emotion_b200 import windowing
    mods = ["video", "vggish", "bert"]
    m = _lfan(mods, dev, seed=2)
    sd = synthetic.lfan_state_dict(2, mods)
    L = 501
    vid = synthetic.frames(L, seed=51)
    g = torch.Generator().manual_seed(52)
    feats = {"vggish": torch.randn(L, 128, generator=g), "bert": torch.randn(L, 768, generator=g)}
    out = windowing.infer_video(m, vid.to(dev), {k: v.to(dev) for k, v in feats.items()}).cpu()
    assert out.shape == (L, 7)
    # oracle: frames are independent in eval mode, so embeddings may be computed once for the check
    emb = O.ir50_forward(sd, vid, "spatial.visual.backbone.")
    X = {"video": emb.view(1, 1, L, 512), "vggish": feats["vggish"].view(1, 1, L, 128), "bert": feats["bert"].view(1, 1, L, 768)}

    def fwd(chunk):
        return O.head_forward(sd, {k: v.squeeze(1) for k, v in chunk.items()}, mods)
    Xo = {"video": X["video"], "vggish": X["vggish"], "bert": X["bert"]}
    # oracle.windowed_inference indexes [:, :, wd] for non-'video' keys and [:, wd] for 'video'
    Xo["video"] = X["video"].squeeze(1)
    want = O.windowed_inference(lambda c: fwd({"video": c["video"].unsqueeze(1), "vggish": c["vggish"], "bert": c["bert"]}), Xo)[0]
    assert (out - want).abs().max().item() <= 2e-2
    assert (out.argmax(-1) == want.argmax(-1)).float().mean().item() >= 0.995


def test_vggish_module_vs_golden(golden_dir):
    """AudioBackbone / VGGish drop-in: bf16 tensor-core convs vs the reference's fp32 output."""
    dev = _dev()
    from feature_vs_text_compound_emotion_b200.models.backbone import AudioBackbone
    g = torch.load(os.path.join(golden_dir, "vggish_n6.pt"))
    ab = AudioBackbone()
    ab.backbone.load_state_dict(synthetic.vggish_state_dict(g["weights_seed"]), strict=True)
    ab = ab.to(dev).eval()
    emb = ab(synthetic.logmel_patches(g["n"], seed=g["x_seed"]).to(dev)).cpu()
    assert emb.shape == (6, 128)
    assert F.cosine_similarity(emb, g["emb"], dim=1).min().item() >= 0.999
    assert ((emb - g["emb"]).abs().max() / g["emb"].abs().max()).item() <= 2e-2


def test_vggish_many_patches_multi_pass():
    """More patches than one pass holds (ragged last pass), against the oracle."""
    dev = _dev()
    from feature_vs_text_compound_emotion_b200.models.backbone import VGGish
    sd = synthetic.vggish_state_dict(4)
    VGGish.patches_per_pass = 40
    try:
        vg = VGGish()
        vg.load_state_dict(sd, strict=True)
        vg = vg.to(dev).eval()
        x = synthetic.logmel_patches(97, seed=11)
        emb = vg(x.to(dev)).cpu()
    finally:
        VGGish.patches_per_pass = 1200
    ref = O.vggish_forward(sd, x)
    assert F.cosine_similarity(emb, ref, dim=1).min().item() >= 0.999


def test_lfan_logmel_from_pixels_vs_golden(golden_dir):
    dev = _dev()
    g = torch.load(os.path.join(golden_dir, "lfan_logmel_b1.pt"))
    mods, T = g["modalities"], g["length"]
    m = _lfan(mods, dev, g["weights_seed"], length=T)
    X = {"video": synthetic.frames(T, seed=g["frame_seed"]).view(1, T, 3, 40, 40).to(dev),
         "logmel": synthetic.logmel_patches(T, seed=g["logmel_seed"]).view(1, T, 96, 64).permute(0, 3, 1, 2).contiguous().to(dev),
         "bert": synthetic.feature_windows(1, T, seed=g["bert_seed"], modalities=["bert"])["bert"].to(dev)}
    out = m(X).cpu()
    assert out.shape == (1, T, 7)
    assert (out - g["logits"]).abs().max().item() <= 2e-2
    assert (out.argmax(-1) == g["logits"].argmax(-1)).float().mean().item() >= 0.975   # 40 frames: at most one flip


def test_infer_video_from_stored_uint8_crops_and_logmel():
    """cfg-3 shape of the path on one short video: stored uint8 256x256 crops -> device eval
    transform -> IR-50; log-mel examples -> VGGish; BERT features; windows; stitch; video vote."""
    dev = _dev()
    from feature_vs_text_compound_emotion_b200 import windowing
    mods = ["video", "logmel", "bert"]
    m = _lfan(mods, dev, seed=6)
    sd = synthetic.lfan_state_dict(6, mods)
    L = 330                                    # two windows: [0,300) and the tail [30,330)
    raw = synthetic.raw_frames_u8(L, seed=61)
    lm = synthetic.logmel_patches(L, seed=62)
    bert = torch.randn(L, 768, generator=torch.Generator().manual_seed(63))
    out = windowing.infer_video(m, raw.to(dev), {"logmel": lm.to(dev), "bert": bert.to(dev)}).cpu()
    assert out.shape == (L, 7)
    vid = O.eval_transform(raw.numpy())
    emb = O.ir50_forward(sd, vid, "spatial.visual.backbone.")
    aud = O.vggish_forward(sd, lm, "spatial.audio.backbone.")
    feats = {"video": emb, "logmel": aud, "bert": bert}
    want = torch.zeros(L, 7)
    cnt = torch.zeros(L, 1)
    for wd in O.windowing(L, 300, 200):
        y = O.head_forward(sd, {k: v[wd].unsqueeze(0) for k, v in feats.items()}, mods)[0]
        want[wd] += y
        cnt[wd] += 1
    want = want / cnt
    assert (out - want).abs().max().item() <= 2e-2
    assert (out.argmax(-1) == want.argmax(-1)).float().mean().item() >= 0.99
    assert windowing.video_level_prediction(out.to(dev))["FRAMES_AVG_LOGITS"] == O.video_level_prediction(want.numpy())["FRAMES_AVG_LOGITS"]


def test_infer_videos_equals_per_video_inference(monkeypatch):
    """windowing.infer_videos: the frames of several videos through the backbones in shared passes and their
    windows through the head in fixed-size groups (a video shorter than one window, exactly one window, grid +
    tail window; a padded last head group).  With one kernel plan for every pass size (CER_RASTER=0) the result
    is bit for bit what infer_video returns per video; with the default plan a pass of >= 190 frames runs IR-50
    stage 2 on the padded-raster kernel, whose fp32 accumulation order differs, so a short video inferred alone
    and inside a batch agree to bf16 rounding instead."""
    dev = _dev()
    from feature_vs_text_compound_emotion_b200 import windowing
    mods = ["video", "logmel", "bert"]
    lengths = [120, 300, 530, 301]
    raw = synthetic.raw_frames_u8(64, seed=71).repeat(9, 1, 1, 1).to(dev)
    lm = synthetic.logmel_patches(max(lengths), seed=72).to(dev)
    bert = torch.randn(max(lengths), 768, generator=torch.Generator().manual_seed(73)).to(dev)
    vids = [raw[7 * i:7 * i + t] for i, t in enumerate(lengths)]
    feats = [{"logmel": lm[:t], "bert": bert[3 * i:3 * i + t] if 3 * i + t <= bert.shape[0] else bert[:t]} for i, t in enumerate(lengths)]

    monkeypatch.setenv("CER_RASTER", "0")
    m = _lfan(mods, dev, seed=6)
    single = [windowing.infer_video(m, v, dict(f)) for v, f in zip(vids, feats)]
    for wpp in (16, 4):                                   # 8 windows in total: one padded group / two full groups
        batch = windowing.infer_videos(m, vids, [dict(f) for f in feats], windows_per_pass=wpp)
        assert [tuple(o.shape) for o in batch] == [(t, 7) for t in lengths]
        for a, b in zip(single, batch):
            assert torch.equal(a, b)
    assert windowing.infer_videos(m, [], []) == []

    monkeypatch.delenv("CER_RASTER")
    m2 = _lfan(mods, dev, seed=6)                         # default plan: engines are built at the first forward
    batch2 = windowing.infer_videos(m2, vids, [dict(f) for f in feats])
    single2 = [windowing.infer_video(m2, v, dict(f)) for v, f in zip(vids, feats)]
    for a, b, c in zip(single, batch2, single2):
        assert (a - b).abs().max().item() <= 1e-2 and (a - c).abs().max().item() <= 1e-2
        assert (a.argmax(-1) == b.argmax(-1)).float().mean().item() >= 0.99


def test_host_prefetcher_matches_direct_calls():
    """Double-buffered H2D staging must not change results or reorder batches."""
    dev = _dev()
    from feature_vs_text_compound_emotion_b200.pipeline import HostPrefetcher
    mods = ["cnn_res50", "vggish", "bert"]
    m = _lfan(mods, dev)
    host = [{k: v.pin_memory() for k, v in synthetic.feature_windows(2, 300, seed=70 + i, modalities=mods).items()} for i in range(5)]
    want = [m({k: v.to(dev) for k, v in h.items()}).cpu() for h in host]
    got = {}
    n = HostPrefetcher(dev).run(iter(host), lambda b: m(b), lambda i, out: got.__setitem__(i, out.cpu()))
    assert n == 5
    for i in range(5):
        assert torch.equal(got[i], want[i]), i


def test_empty_ragged_and_short_inputs():
    """Edge cases of the entry points: zero frames, a ragged second pass, a video shorter than one
    window (padded by repeating its last frame, base/dataset.py:570-582), a single-frame video."""
    dev = _dev()
    from feature_vs_text_compound_emotion_b200 import windowing
    from feature_vs_text_compound_emotion_b200.engine import PreprocEngine
    from feature_vs_text_compound_emotion_b200.models.arcface_model import Backbone
    mods = ["video", "vggish", "bert"]
    m = _lfan(mods, dev, seed=7)
    sd = synthetic.lfan_state_dict(7, mods)
    vb = m.spatial["visual"]
    assert vb(torch.empty(0, 3, 40, 40, device=dev)).shape == (0, 512)
    assert PreprocEngine(256, 256, dev).forward(torch.empty(0, 256, 256, 3, dtype=torch.uint8, device=dev)).shape == (0, 3, 40, 40)
    # ragged passes: 2 full passes of 16 frames + 1 frame
    old = Backbone.frames_per_pass
    Backbone.frames_per_pass = 16
    try:
        vb.backbone.repack()
        x = synthetic.frames(33, seed=71)
        emb = vb(x.to(dev)).cpu()
    finally:
        Backbone.frames_per_pass = old
        vb.backbone.repack()
    ref = O.ir50_forward(sd, x, "spatial.visual.backbone.")
    assert F.cosine_similarity(emb, ref, dim=1).min().item() >= 0.999
    # short videos
    for L in (1, 7, 299):
        vid = synthetic.frames(L, seed=72 + L)
        g = torch.Generator().manual_seed(73 + L)
        feats = {"vggish": torch.randn(L, 128, generator=g), "bert": torch.randn(L, 768, generator=g)}
        out = windowing.infer_video(m, vid.to(dev), {k: v.to(dev) for k, v in feats.items()}).cpu()
        assert out.shape == (L, 7)
        emb = O.ir50_forward(sd, vid, "spatial.visual.backbone.")
        pad = lambda t: torch.cat([t, t[-1:].expand(300 - L, -1)]).unsqueeze(0)
        want = O.head_forward(sd, {"video": pad(emb), "vggish": pad(feats["vggish"]), "bert": pad(feats["bert"])}, mods)[0, :L]
        assert (out - want).abs().max().item() <= 2e-2, L


@pytest.mark.parametrize("name,fixture,bar", [("CAN", "heads.pt", 0.95), ("JMT", "heads.pt", 0.95), ("MT", "heads.pt", 0.95),
                                              ("CAN", "heads_t300.pt", 0.995), ("JMT", "heads_t300.pt", 0.995),
                                              ("MT", "heads_t300.pt", 0.995)])
def test_alternative_heads_vs_golden(golden_dir, name, fixture, bar):
    """CAN / JMT / MT drop-ins from pixels vs the reference's outputs: B = 2 windows of T = 24 frames (48
    frames: at most two argmax flips) and of the reference's window length T = 300, where JMT / MT run
    nn.MultiheadAttention(128, 1) over 300 positions and the final encoder over all 600 (models/model.py:731,
    :917-931, :1003-1012) -- BASELINE.json's bars: logit max-abs <= 2e-2, argmax agreement >= 99.5 %."""
    dev = _dev()
    from feature_vs_text_compound_emotion_b200.models.model import CAN, JMT
    g = torch.load(os.path.join(golden_dir, fixture))[name]
    mods, T = g["modalities"], g["T"]
    vsd = synthetic.visual_backbone_state_dict(0)
    if name == "CAN":
        m = CAN(task="CLASSIFICATION", modalities=mods, tcn_settings=synthetic.TCN_SETTINGS, backbone_settings=BS, output_dim=7,
                root_dir="", device=dev, visual_state_dict=vsd)
        m.load_state_dict(synthetic.can_state_dict(0, mods), strict=True)
    else:
        m = JMT(task="CLASSIFICATION", modalities=mods, tcn_settings=synthetic.TCN_SETTINGS, backbone_settings=BS, output_dim=7,
                root_dir="", device=dev, model_name=name, visual_state_dict=vsd)
        m.load_state_dict(synthetic.jmt_state_dict(0, mods, model_name=name), strict=True)
    m = m.to(dev).eval()
    X = {"video": synthetic.frames(2 * T, seed=g["frame_seed"]).view(2, T, 3, 40, 40).to(dev)}
    for k, v in synthetic.feature_windows(2, T, seed=g["feat_seed"], modalities=[x for x in mods if x != "video"]).items():
        X[k] = v.to(dev)
    out = m(X).cpu()
    assert out.shape == g["out"].shape
    err = (out - g["out"]).abs().max().item()
    agree = (out.argmax(-1) == g["out"].argmax(-1)).float().mean().item()
    assert err <= 2e-2 * max(1.0, g["out"].abs().max().item()), err
    assert agree >= bar, agree


def test_vggish_fused_pool_equals_separate_pool_kernel(monkeypatch):
    """The max-pool fused into the conv epilogue (shuffles over bf16-rounded values) must be
    bit-identical to pooling the stored bf16 tensor with the stand-alone kernel."""
    dev = _dev()
    from feature_vs_text_compound_emotion_b200 import packing
    from feature_vs_text_compound_emotion_b200.engine import VggishEngine
    pk = packing.pack_vggish(synthetic.vggish_state_dict(2))
    x = synthetic.logmel_patches(45, seed=21).to(dev)
    fused = VggishEngine(pk, dev, patches_per_pass=64).forward(x).cpu()
    monkeypatch.setenv("CER_NO_POOL_FUSION", "1")
    plain_eng = VggishEngine(pk, dev, patches_per_pass=64)
    assert plain_eng.launches(45) == 12 and True
    plain = plain_eng.forward(x).cpu()
    assert torch.equal(fused, plain)


def test_get_attention_maps_vs_golden(golden_dir):
    dev = _dev()
    g = torch.load(os.path.join(golden_dir, "attention_maps.pt"))
    mods = ["video", "vggish", "bert"]
    m = _lfan(mods, dev)
    x = {k: torch.randn(2, g["T"], d, generator=torch.Generator().manual_seed(g["seed"])).to(dev) for k, d in zip(mods, (128, 32, 128))}
    maps = m.fusion.get_attention_maps(x)
    assert isinstance(maps, list) and len(maps) == 1 and maps[0].shape == (2, 2, g["T"], 3, 3)
    assert (maps[0].cpu() - g["maps"]).abs().max().item() < 1e-5
    assert torch.allclose(maps[0].sum(-1).cpu(), torch.ones(2, 2, g["T"], 3), atol=1e-5)


def test_full_size_properties_config2():
    """BASELINE configs[1] size (8 x 300 frames): properties that need no oracle run at that size --
    unit-norm embeddings, bit-exact determinism, frame-permutation equivariance (every frame is
    independent in eval mode, whatever tile it lands in), pass-size invariance, and causality /
    window independence of the head."""
    dev = _dev()
    from feature_vs_text_compound_emotion_b200.models.arcface_model import Backbone
    mods = ["video", "vggish", "bert"]
    m = _lfan(mods, dev, seed=8)
    vb = m.spatial["visual"]
    n = 2400
    x = synthetic.frames(n, seed=81).to(dev)
    e1 = vb(x)
    assert torch.allclose(e1.norm(dim=1), torch.ones(n, device=dev), atol=1e-4)
    assert torch.equal(vb(x), e1)                                        # deterministic
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(82)).to(dev)
    assert torch.equal(vb(x[perm]), e1[perm])                            # frames are independent
    old = Backbone.frames_per_pass
    Backbone.frames_per_pass = 700                                       # 3 full passes + a ragged one
    try:
        vb.backbone.repack()
        e2 = vb(x)
    finally:
        Backbone.frames_per_pass = old
        vb.backbone.repack()
    assert torch.equal(e2, e1)
    # head: logits of frame t depend only on frames <= t of the same window
    f = synthetic.feature_windows(8, 300, seed=83, modalities=["vggish", "bert"])
    X = {"video": e1.view(8, 300, 512), "vggish": f["vggish"].squeeze(1).to(dev), "bert": f["bert"].squeeze(1).to(dev)}
    y0 = m.forward_features({k: v.clone() for k, v in X.items()})
    X2 = {k: v.clone() for k, v in X.items()}
    X2["bert"][:, 200:] += 1.0                                           # change the future of every window
    X2["vggish"][3] = 0.0                                                # and all of window 3
    y1 = m.forward_features(X2)
    keep = [w for w in range(8) if w != 3]
    assert torch.equal(y1[keep, :200], y0[keep, :200])
    assert not torch.equal(y1[keep, 200:], y0[keep, 200:])
    assert not torch.equal(y1[3], y0[3])


def test_lfan_variants_modalities_heads_regression():
    """Other configurations the constructor allows: two modalities with another leader, 4 heads of 16,
    REGRESSION task (tanh on the output, model.py:523) with 2 outputs."""
    dev = _dev()
    from feature_vs_text_compound_emotion_b200.models.model import LFAN
    mods = ["vggish", "bert"]
    kw = dict(modal_dim=64, tcn_channels=synthetic.TCN_CHANNELS)
    sd = synthetic.head_state_dict(9, mods, output_dim=2, modal_dim=64)
    m = LFAN(backbone_settings=BS, output_dim=2, task="REGRESSION", modality=mods, kernel_size=5, example_length=300,
             tcn_channel=synthetic.TCN_CHANNELS, modal_dim=64, num_heads=4, root_dir="", device=dev)
    m.init()
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    X = synthetic.feature_windows(2, 300, seed=91, modalities=mods)
    want = torch.tanh(O.lfan_forward(sd, {k: v.clone() for k, v in X.items()}, mods, modal_dim=64, num_heads=4))
    out = m({k: v.to(dev) for k, v in X.items()}).cpu()
    assert out.shape == (2, 300, 2)
    assert (out - want).abs().max().item() <= 5e-3
    del kw


def test_lfan_odd_feature_widths_mfcc_egemaps():
    """mfcc (39-d) and egemaps (88-d) inputs (model.py:388-393) are not multiples of the tensor-core
    kernel's 32-channel chunks: weights and features are zero-padded, results unchanged."""
    dev = _dev()
    from feature_vs_text_compound_emotion_b200.models.model import LFAN
    mods = ["mfcc", "egemaps", "bert"]
    sd = synthetic.head_state_dict(10, mods)
    m = LFAN(backbone_settings=BS, output_dim=7, task="CLASSIFICATION", modality=mods, kernel_size=5, example_length=300,
             tcn_channel=synthetic.TCN_CHANNELS, modal_dim=32, num_heads=2, root_dir="", device=dev)
    m.init()
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    X = synthetic.feature_windows(2, 300, seed=92, modalities=mods)
    want = O.lfan_forward(sd, {k: v.clone() for k, v in X.items()}, mods)
    out = m({k: v.to(dev) for k, v in X.items()}).cpu()
    assert (out - want).abs().max().item() <= 5e-3
    assert (out.argmax(-1) == want.argmax(-1)).float().mean().item() >= 0.995


def test_backbone_standalone_7x7_head_56px():
    """arcface_model.Backbone(50, ...) as the reference constructs it stand-alone: 7x7 output layer,
    i.e. 56x56 inputs with this repo's stage-1 stride (arcface_model.py:98, :133-137)."""
    dev = _dev()
    from feature_vs_text_compound_emotion_b200.models.arcface_model import Backbone
    sd = {k[len("backbone."):]: v for k, v in synthetic.visual_backbone_state_dict(11, spatial=7).items() if k.startswith("backbone.")}
    bb = Backbone(50, 0.4, 3, 'ir')
    bb.load_state_dict(sd, strict=True)
    bb = bb.to(dev).eval()
    x = synthetic.frames(6, seed=93, size=56)
    emb = bb(x.to(dev)).cpu()
    ref = O.ir50_forward(sd, x, "")
    assert emb.shape == (6, 512)
    assert F.cosine_similarity(emb, ref, dim=1).min().item() >= 0.999


def test_infer_video_from_waveform():
    """The audio branch from the raw 16 kHz waveform: edge padding (vggish_input.py:93), log-mel,
    one 0.96 s example per video frame, VGGish, then the usual path."""
    dev = _dev()
    import numpy as np
    from feature_vs_text_compound_emotion_b200 import windowing
    mods = ["video", "logmel", "bert"]
    m = _lfan(mods, dev, seed=12)
    sd = synthetic.lfan_state_dict(12, mods)
    L, fps = 45, 30.0
    wave = synthetic.waveform(L / fps, seed=121)                 # exactly as long as the video: needs the padding
    vid = synthetic.frames(L, seed=122)
    bert = torch.randn(L, 768, generator=torch.Generator().manual_seed(123))
    out = windowing.infer_video(m, vid.to(dev), {"wave": wave.to(dev), "bert": bert.to(dev)}, fps=fps).cpu()
    assert out.shape == (L, 7)
    w = wave.double().numpy()
    w = np.pad(w, (0, 16000), "edge")
    ex = torch.from_numpy(O.waveform_to_examples(w, 0.96, 1.0 / fps)).float()
    assert ex.shape[0] >= L
    aud = O.vggish_forward(sd, ex[:L], "spatial.audio.backbone.")
    emb = O.ir50_forward(sd, vid, "spatial.visual.backbone.")
    pad = lambda t: torch.cat([t, t[-1:].expand(300 - L, -1)]).unsqueeze(0)
    want = O.head_forward(sd, {"video": pad(emb), "logmel": pad(aud), "bert": pad(bert)}, mods)[0, :L]
    assert (out - want).abs().max().item() <= 2e-2


def test_video_level_decision_rules_match_reference_semantics():
    """cer_video_vote vs the restated format_trg_pred_video (metrics.py:118-142), incl. the tie rule."""
    dev = _dev()
    from feature_vs_text_compound_emotion_b200 import windowing
    gen = torch.Generator().manual_seed(9)
    for T in (1, 2, 7, 300, 1001, 5000):
        lg = torch.randn(T, 8, generator=gen)
        for ign in (False, True):
            assert windowing.video_level_prediction(lg.to(dev), ign) == O.video_level_prediction(lg.numpy(), ign), (T, ign)
    # tie: classes 2 and 5 both win 2 frames, class 5 appears first -> Counter.most_common picks 5
    lg = torch.full((4, 7), -1.0)
    for t, c in enumerate((5, 2, 2, 5)):
        lg[t, c] = 1.0
    assert windowing.video_level_prediction(lg.to(dev))["FRAMES_VOTE"] == 5 == O.video_level_prediction(lg.numpy())["FRAMES_VOTE"]


def test_full_size_oracle_parity_config2():
    """The BENCHMARKED size and plan: 8 windows x 300 frames from pixels through LFAN.forward with the
    default frames_per_pass (so the CTA-pair / halo / resident-weight kernel selection bench.py times is
    the one checked) against the oracle's fp32 CPU forward (models/model.py:487-526 at B=8, T=300).
    BASELINE.json bars on every one of the 2400 frames."""
    dev = _dev()
    from feature_vs_text_compound_emotion_b200.models.arcface_model import Backbone
    assert Backbone.frames_per_pass == 2400
    mods = ["video", "vggish", "bert"]
    m = _lfan(mods, dev, seed=0)
    sd = synthetic.lfan_state_dict(0, mods)
    feats = synthetic.feature_windows(8, 300, seed=101, modalities=["vggish", "bert"])
    vid = synthetic.frames(2400, seed=102).view(8, 300, 3, 40, 40)
    emb = m.spatial["visual"](vid.view(2400, 3, 40, 40).to(dev)).cpu()
    ref_emb = O.ir50_forward(sd, vid.view(2400, 3, 40, 40), "spatial.visual.backbone.")
    cos = F.cosine_similarity(emb, ref_emb, dim=1)
    assert cos.min().item() >= 0.999, cos.min().item()
    X = {"video": vid, "vggish": feats["vggish"], "bert": feats["bert"]}
    out = m({k: v.to(dev) for k, v in X.items()}).cpu()
    ref = O.head_forward(sd, {"video": ref_emb.view(8, 300, 512), "vggish": feats["vggish"].squeeze(1),
                              "bert": feats["bert"].squeeze(1)}, mods)
    assert out.shape == ref.shape == (8, 300, 7)
    err = (out - ref).abs().max().item()
    agree = (out.argmax(-1) == ref.argmax(-1)).float().mean().item()
    assert err <= 2e-2, err
    assert agree >= 0.995, agree


def test_head_graph_replay_after_workspace_regrow():
    """A CUDA graph captured for a small (B, T) must stay valid after a larger input made the TCN
    engines grow their workspace (the graph's kernel nodes hold the old block's address)."""
    dev = _dev()
    mods = ["cnn_res50", "vggish", "bert"]
    m = _lfan(mods, dev, seed=13)
    small = {k: v.squeeze(1).to(dev) for k, v in synthetic.feature_windows(1, 300, seed=131, modalities=mods).items()}
    big = {k: v.squeeze(1).to(dev) for k, v in synthetic.feature_windows(6, 300, seed=132, modalities=mods).items()}
    y_small = m.forward_features(dict(small)).clone()          # captures the (1, 300) graph
    m.forward_features(dict(small))
    keep = [m.forward_features(dict(big)).clone() for _ in range(2)]      # regrows the workspaces
    junk = [torch.full((1 << 20,), float("nan"), device=dev) for _ in range(8)]   # would land in a freed block
    y_again = m.forward_features(dict(small))                  # replays the old graph
    assert torch.equal(y_again, y_small)
    m.head_cuda_graph = False
    try:
        assert torch.equal(m.forward_features(dict(small)), y_small)
        assert torch.equal(m.forward_features(dict(big)), keep[0])
    finally:
        m.head_cuda_graph = True
    assert all(torch.isnan(j).all() for j in junk)
