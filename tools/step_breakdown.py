"""Where one LFAN step (8 x 300 frames) spends its time: CUDA events around IR-50, each TCN and the
fusion kernel inside the real forward, vs the whole step (back-to-back steps)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch
import bench

torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
model = bench.build_model(dev)
batch = {k: v.to(dev) for k, v in bench.host_batch(100).items()}
reps = 20
for _ in range(3):
    model(dict(batch))
torch.cuda.synchronize()
ev = lambda: torch.cuda.Event(enable_timing=True)
acc = {}
def seg(name, fn):
    a, b = ev(), ev()
    a.record(); out = fn(); b.record()
    acc.setdefault(name, []).append((a, b))
    return out
t0, t1 = ev(), ev()
t0.record()
for _ in range(reps):
    vid = batch["video"].view(2400, 3, 40, 40)
    emb = seg("ir50", lambda: model.spatial["visual"](vid)).view(8, 300, 512)
    tcn, fus = model._head_engines()
    enc = []
    for m, x in (("video", emb), ("vggish", batch["vggish"].squeeze(1)), ("bert", batch["bert"].squeeze(1))):
        enc.append(seg("tcn_" + m, lambda: tcn[m].forward(x)))
    seg("fusion", lambda: fus.forward([e.view(2400, -1) for e in enc]))
t1.record()
torch.cuda.synchronize()
tot = t0.elapsed_time(t1) / reps
print(f"segmented step: {tot:.3f} ms")
for k, v in acc.items():
    print(f"  {k:12s} {sum(a.elapsed_time(b) for a, b in v) / reps:.3f} ms")
s0, s1 = ev(), ev()
s0.record()
for _ in range(reps):
    model(dict(batch))
s1.record()
torch.cuda.synchronize()
print(f"model(X) back to back: {s0.elapsed_time(s1) / reps:.3f} ms")
