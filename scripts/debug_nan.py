"""Development aid: run the SD-2-size C1 forward with every GEMM / attention output checked for NaN, print the first offender."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from panfusion_b200 import ops, sd2_unet
from panfusion_b200.mvgen import MultiViewBaseModel

dev = torch.device("cuda:0")
dt = torch.float16
model = MultiViewBaseModel(sd2_unet.build_synthetic(seed=1, device=dev), sd2_unet.build_synthetic(seed=2, device=dev),
                           compute_dtype=dt, overlap_branches=False).to(dev).eval()
g = torch.Generator(device=dev).manual_seed(4)
with torch.no_grad():
    for name, p in sorted(model.named_parameters()):
        if "cp_blocks" in name and float(p.abs().sum()) == 0.0:
            p.copy_(torch.randn(p.shape, generator=g, device=dev) * 0.02)
model.prepare(dev, dt)
orig = ops.gemm_taps
count = [0]
def wrapped(A, B, out, **kw):
    r = orig(A, B, out, **kw)
    torch.cuda.synchronize()
    o = r[0] if isinstance(r, tuple) else r
    bad = bool(torch.isnan(o.float()).any()) or bool(torch.isinf(o.float()).any())
    sbad = isinstance(r, tuple) and (bool(torch.isnan(r[1]).any()) or bool(torch.isinf(r[1]).any()))
    count[0] += 1
    if bad or sbad or count[0] <= 0:
        ln = kw.get("ln")
        print(f"call {count[0]}: M={kw['M']} N={B.shape[0]} Kc={kw['Kc']} taps={len(kw.get('taps', (0,)))} act={kw.get('act', 0)} "
              f"row_stats={kw.get('row_stats', False)} ln={'yes slots=%d' % ln[0].shape[1] if ln else 'no'} block_n={kw.get('block_n', 0)} "
              f"out_nan={bad} stats_nan={sbad} A_nan={bool(torch.isnan(A.float()).any())}", flush=True)
        if ln:
            st = ln[0]
            print("  ln stats nan:", bool(torch.isnan(st).any()), "min sumsq", float(st[..., 1].min()), "colsum nan", bool(torch.isnan(ln[1]).any()),
                  "stats shape", tuple(st.shape), flush=True)
        raise SystemExit(1)
    return r
ops.gemm_taps = wrapped
import panfusion_b200.engine as E, panfusion_b200.eppa as P
m = 2
gg = torch.Generator().manual_seed(0)
import numpy as np
theta = torch.tensor(np.rad2deg(np.linspace(0, 2 * np.pi, m, endpoint=False)), dtype=torch.float32)[None]
cams = dict(FoV=torch.full((1, m), 90.0).to(dev), theta=theta.to(dev), phi=torch.zeros(1, m).to(dev))
pano = torch.randn(1, 1, 4, 64, 128, generator=gg).to(dev)
lat = torch.randn(1, m, 4, 64, 64, generator=gg).to(dev)
prompt = torch.randn(1, m, 77, 1024, generator=gg).to(dev)
pp = torch.randn(1, 1, 77, 1024, generator=gg).to(dev)
ts = torch.full((1, m), 981, dtype=torch.long, device=dev)
s, p = model(lat, pano, ts, prompt, pp, cams)
print("ok", count[0], bool(torch.isnan(s).any()), bool(torch.isnan(p).any()))
