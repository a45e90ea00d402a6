"""Time (and let ncu profile) one tap-GEMM shape: python scripts/gemm_micro.py M N Kc taps [block_n] [act] [res]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from panfusion_b200 import ops  # noqa: E402

M, N, Kc, ntaps = (int(v) for v in sys.argv[1:5])
bn = int(sys.argv[5]) if len(sys.argv) > 5 else 0
act = int(sys.argv[6]) if len(sys.argv) > 6 else 0
has_res = int(sys.argv[7]) if len(sys.argv) > 7 else 0
mapped = int(sys.argv[8]) if len(sys.argv) > 8 else 0
dev = torch.device("cuda:0")
A = torch.randn(M + 4096, Kc, device=dev).bfloat16()
B = (torch.randn(N, Kc * ntaps, device=dev) * 0.02).bfloat16()
n_out = N // 2 if act == 3 else N
out = torch.empty(M, n_out, dtype=torch.bfloat16, device=dev)
res = torch.randn(M, n_out, device=dev).bfloat16() if has_res else None
fn = lambda: ops.gemm_taps(A, B, out, M=M, Kc=Kc, taps=list(range(ntaps)), residual=res, act=act, block_n=bn,
                           image_map=(1, M, 0, 0, 1, M) if mapped else None)
for _ in range(3):
    fn()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    fn()
b.record()
torch.cuda.synchronize()
us = a.elapsed_time(b) / 20 * 1e3
fl = 2.0 * M * N * Kc * ntaps
print(f"M={M} N={N} Kc={Kc} taps={ntaps} bn={bn or 'auto'} act={act} res={has_res}: {us:.1f} us, {fl / us / 1e6:.1f} TFLOP/s")
