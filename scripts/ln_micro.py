"""Time the linear-layer GEMM with and without the fused-LayerNorm producer / consumer epilogues (development aid)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from panfusion_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for (M, N, K) in ((16384, 320, 320), (65536, 320, 320), (4096, 640, 640), (16384, 960, 320), (2048, 320, 320)):
    A = torch.randn(M, K, device=dev).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    res = torch.randn(M, N, device=dev).bfloat16()
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    bias = torch.randn(N, device=dev)
    colsum = torch.randn(N, device=dev)
    _, st = ops.gemm_taps(A, B, out, M=M, Kc=K, bias=bias, residual=res if N == K else None, row_stats=True)
    if N == K:
        t0 = timeit(lambda: ops.gemm_taps(A, B, out, M=M, Kc=K, bias=bias, residual=res))
        t1 = timeit(lambda: ops.gemm_taps(A, B, out, M=M, Kc=K, bias=bias, residual=res, row_stats=True))
    else:
        t0 = t1 = float("nan")
    st_k = torch.randn(M, 2 * max(1, K // 160), 2, device=dev).abs() + 1
    t2 = timeit(lambda: ops.gemm_taps(A, B, out, M=M, Kc=K, bias=bias))
    t3 = timeit(lambda: ops.gemm_taps(A, B, out, M=M, Kc=K, bias=bias, ln=(st_k, colsum, 1e-5)))
    print(f"M={M} N={N} K={K}: plain+res {t0:.1f} us | producer(row_stats)+res {t1:.1f} us | plain {t2:.1f} us | consumer(ln) {t3:.1f} us",
          flush=True)
