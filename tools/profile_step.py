"""One LFAN forward (BASELINE configs[1] shape) after two warm-up forwards -- the command ncu
wraps.  77 kernel launches per forward: stem, 48 unit convs (+ raster_zero_pads_kernel in a raster pass), FC,
l2norm, 24 TCN launches (two per TemporalBlock), fusion."""
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import torch  # noqa: E402

import bench  # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
model = bench.build_model(dev)
batch = {k: v.to(dev) for k, v in bench.host_batch(100).items()}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(n):
    out = model(dict(batch))
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out.abs().mean()))
