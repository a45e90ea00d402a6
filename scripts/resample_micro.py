"""Time (and let ncu profile) the e2p kernel at the reference's hot-path shape: fp32 (16,2048,32,64) -> (16,2048,32,32)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from panfusion_b200 import geometry  # noqa: E402

dev = torch.device("cuda:0")
x = torch.randn(16, 2048, 32, 64, device=dev)
th = torch.tensor(np.tile(np.arange(8) * 45.0, 2), dtype=torch.float32)
fov, phi = torch.full((16,), 90.0), torch.zeros(16)
fn = lambda: geometry.e2p(x, fov, th, phi, (32, 32))
for _ in range(3):
    fn()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    fn()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
alg = x.numel() * 4 + 16 * 2048 * 32 * 32 * 4
print(f"e2p fp32 16x2048x32x64 -> 32x32: {ms * 1e3:.1f} us, {alg / ms / 1e6:.1f} GB/s algorithmic")
