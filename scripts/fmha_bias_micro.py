"""Graph-timed d=32 EPPA attention at the C2 level-32 shapes with the real tile-packed correspondence bias (direction 1: 2048
panorama queries x 8192 view keys; direction 2: 8192 view queries x 2048 panorama keys), next to the same launch with every
tile constant (-1) and with no bias: what the bias path costs. PF_LIB_PATH selects the build (A/B)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from panfusion_b200 import ops  # noqa: E402
from panfusion_b200.eppa import CameraTables  # noqa: E402

dev, dt = torch.device("cuda:0"), torch.bfloat16
torch.manual_seed(0)
m, b = 8, 2
heads = int(sys.argv[1]) if len(sys.argv) > 1 else 10
c = heads * 32
theta = torch.tensor(np.tile(np.arange(m) * 45.0, b), dtype=torch.float32)
cams = dict(FoV=torch.full((b * m,), 90.0), theta=theta, phi=torch.zeros(b * m))
key = CameraTables.camera_key(cams)
tables = CameraTables()
(d1, d2) = tables.bias(*CameraTables.dedup(key, b), 32, 32, 32, 64, dev)
E, P = 32 * 64, m * 32 * 32


def timed(fn, reps=20):
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn()
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps):
                fn()
    g.replay()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return a.elapsed_time(e) / (5 * reps) * 1e3


for name, (Lq, Lk, d) in dict(dir1=(E, P, d1), dir2=(P, E, d2)).items():
    q = torch.randn(b, Lq, c, device=dev).to(dt)
    k = torch.randn(b, Lk, c, device=dev).to(dt)
    v = torch.randn(b, Lk, c, device=dev).to(dt)
    o = torch.empty_like(q)
    store, off = d[1], d[2]
    const_off = torch.full_like(off, -1)
    t_real = timed(lambda: ops.fmha(q, k, v, o, heads=heads, head_dim=32, scale=32 ** -0.5, bias_tiles=(store, off)))
    t_const = timed(lambda: ops.fmha(q, k, v, o, heads=heads, head_dim=32, scale=32 ** -0.5, bias_tiles=(store, const_off)))
    t_none = timed(lambda: ops.fmha(q, k, v, o, heads=heads, head_dim=32, scale=32 ** -0.5))
    live = int((off >= 0).sum())
    print(f"{name} H={heads} Lq={Lq} Lk={Lk}: packed bias {t_real:.1f} us ({live}/{off.numel()} tiles live), all tiles constant "
          f"{t_const:.1f} us, no bias {t_none:.1f} us")
