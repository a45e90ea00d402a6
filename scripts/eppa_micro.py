"""One EPPA fusion block (WarpAttn, models/pano/modules.py:15-59) at the C2 level-32 shape through the public module, for
`ncu -k regex:fmha` (the d=32 attention with the tile-packed correspondence bias) and for a footprint report."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from panfusion_b200.engine import Img  # noqa: E402
from panfusion_b200.eppa import CameraTables, WarpAttn  # noqa: E402

dev, dt = torch.device("cuda:0"), torch.bfloat16
torch.manual_seed(0)
blk = WarpAttn(320).to(dev).eval()
with torch.no_grad():
    for p in blk.parameters():
        if float(p.abs().sum()) == 0.0:
            p.copy_(torch.randn_like(p) * 0.02)
m, b = 8, 2
theta = torch.tensor(np.tile(np.arange(m) * 45.0, b), dtype=torch.float32)
cams = dict(FoV=torch.full((b * m,), 90.0), theta=theta, phi=torch.zeros(b * m))
key = CameraTables.camera_key(cams)
pers = Img(torch.randn(b * m * 32 * 32, 320, device=dev).to(dt), b * m, 32, 32)
equi = Img(torch.randn(b * 32 * 64, 320, device=dev).to(dt), b, 32, 64)
for _ in range(3):
    blk.forward_tokens(pers, equi, key)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    blk.forward_tokens(pers, equi, key)
e.record()
torch.cuda.synchronize()
(d1, d2) = blk.tables.bias(*CameraTables.dedup(key, b), 32, 32, 32, 64, dev)
dense = 2 * (32 * 64) * (m * 32 * 32) * 4
live1, live2 = int((d1[2] >= 0).sum()), int((d2[2] >= 0).sum())
print(f"EPPA block C2 level 32: {a.elapsed_time(e) / 10 * 1e3:.1f} us per fusion; bias tiles live {live1}/{d1[2].numel()} + {live2}/{d2[2].numel()}; "
      f"resident {(d1[1].numel() + d2[1].numel()) * 4 / 2**20:.1f} MB vs dense {dense / 2**20:.1f} MB")
