"""Run ONE launch of a resampling kernel at a benchmark shape (for `ncu --set full`): python scripts/resample_micro.py
{e2p|p2e|e2p_bf16|py360}. Shapes: SURVEY.md 8d (i) e2p fp32 (16,2048,32,64)->(32,32), p2e fp32 (16,1024,32,32)->(32,64),
(ii) e2p bf16 (2,320,64,128)->(16,320,64,64); py360 = 20 views 512x512 of a 1024x2048 uint8 panorama (dataset path)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from panfusion_b200 import geometry, py360  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "e2p"
dev = torch.device("cuda:0")
th = torch.tensor(np.tile(np.arange(8) * 45.0, 2), dtype=torch.float32)
fov, phi = torch.full((16,), 90.0), torch.zeros(16)
if which == "e2p":
    x = torch.randn(16, 2048, 32, 64, device=dev)
    fn = lambda: geometry.e2p(x, fov, th, phi, (32, 32))
elif which == "p2e":
    x = torch.randn(16, 1024, 32, 32, device=dev)
    fn = lambda: geometry.p2e(x, fov, th, phi, (32, 64))
elif which == "e2p_bf16":
    x = torch.randn(2, 320, 64, 128, device=dev).bfloat16()
    fn = lambda: geometry.e2p(x, fov, th, phi, (64, 64), views_per_image=8)
else:
    x = torch.randint(0, 256, (1024, 2048, 3), dtype=torch.uint8, device=dev)
    yaw, pitch = np.linspace(-180, 180, 20, endpoint=False), np.tile([52.6, 10.8, -10.8, -52.6], 5)
    fn = lambda: py360.e2p_views(x, (90, 90), yaw, pitch, (512, 512))
for _ in range(3):
    fn()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    fn()
b.record()
torch.cuda.synchronize()
print(f"{which}: {a.elapsed_time(b) / 10 * 1e3:.1f} us per call (Python launch path included)")
# the same call captured in a CUDA graph (20 launches per replay): kernel time without the host launch path
side = torch.cuda.Stream()
g = torch.cuda.CUDAGraph()
with torch.cuda.stream(side):
    fn()
    side.synchronize()
    with torch.cuda.graph(g, stream=side):
        for _ in range(20):
            fn()
g.replay()
torch.cuda.synchronize()
a.record()
for _ in range(5):
    g.replay()
b.record()
torch.cuda.synchronize()
print(f"{which}: {a.elapsed_time(b) / 100 * 1e3:.1f} us per launch (graph-timed)")
torch.cuda.profiler.start()
fn()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
