"""Time (and let ncu profile) one attention shape: python scripts/fmha_micro.py B H Lq Lk d [bias]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from panfusion_b200 import ops  # noqa: E402

B, H, Lq, Lk, d = (int(v) for v in sys.argv[1:6])
has_bias = int(sys.argv[6]) if len(sys.argv) > 6 else 0
dev = torch.device("cuda:0")
C = H * d
q = torch.randn(B, Lq, C, device=dev).bfloat16()
k = torch.randn(B, Lk, C, device=dev).bfloat16()
v = torch.randn(B, Lk, C, device=dev).bfloat16()
o = torch.empty(B, Lq, C, dtype=torch.bfloat16, device=dev)
bias = (torch.rand(Lq, Lk, device=dev) * 2 - 1) if has_bias else None
fn = lambda: ops.fmha(q, k, v, o, heads=H, head_dim=d, scale=d ** -0.5, bias=bias)
for _ in range(3):
    fn()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    fn()
b.record()
torch.cuda.synchronize()
us = a.elapsed_time(b) / 10 * 1e3
fl = 4.0 * B * H * Lq * Lk * d
print(f"B={B} H={H} Lq={Lq} Lk={Lk} d={d} bias={has_bias}: {us:.1f} us, {fl / us / 1e6:.1f} TFLOP/s")
