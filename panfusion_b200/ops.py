"""Thin Python wrappers over the C-ABI kernels (raw pointers + current CUDA stream). No math happens here."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch
from torch import Tensor

from . import _lib
from ._lib import PF_ACT_GEGLU, PF_ACT_GELU, PF_ACT_NONE, PF_ACT_SILU, FmhaArgs, GemmArgs  # noqa: F401


def _vp(t: Optional[Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _st():
    return C.c_void_p(_lib.stream_ptr())


def pick_block_n(n: int, act: int = PF_ACT_NONE) -> int:
    return int(_lib.lib().pf_gemm_pick_block_n(int(n), int(act)))


def gemm_taps(A: Tensor, B: Tensor, out: Tensor, *, M: int, Kc: int, taps: Sequence[int] = (0,),
              bias: Optional[Tensor] = None, rowbias: Optional[Tensor] = None, rows_per_group: int = 0,
              residual: Optional[Tensor] = None, act: int = PF_ACT_NONE,
              image_map: Optional[tuple] = None, block_n: int = 0) -> Tensor:
    """acc = sum_t A[m + taps[t], :Kc] @ B[:, t*Kc:(t+1)*Kc]^T ; see include/panfusion_b200.h (pf_gemm_taps).

    A: [a_rows, a_ld] 16-bit, B: [N, len(taps)*Kc] 16-bit packed weight, out: [rows, n_out].
    image_map = (Hm, Wm, i0, j0, Hout, Wout) selects map_mode 1.
    """
    _lib.require_cuda(A, B, out)
    assert A.dim() == 2 and B.dim() == 2 and out.dim() == 2
    assert A.stride(1) == 1 and B.stride(1) == 1 and out.stride(1) == 1
    a = GemmArgs()
    a.A, a.a_rows, a.a_ld = A.data_ptr(), A.shape[0], A.stride(0)
    a.B, a.b_ld = B.data_ptr(), B.stride(0)
    a.dtype = _lib.dtype_code(A.dtype)
    assert B.dtype == A.dtype
    a.M, a.N, a.Kc, a.num_taps = int(M), B.shape[0], int(Kc), len(taps)
    for i, t in enumerate(taps):
        a.tap_off[i] = int(t)
    a.block_n = int(block_n)
    a.out, a.out_ld, a.out_dtype = out.data_ptr(), out.stride(0), _lib.dtype_code(out.dtype)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
        a.bias = bias.data_ptr()
    if rowbias is not None:
        assert rowbias.dtype == torch.float32 and rowbias.stride(1) == 1
        a.rowbias, a.rowbias_ld, a.rows_per_group = rowbias.data_ptr(), rowbias.stride(0), int(rows_per_group)
    if residual is not None:
        assert residual.stride(1) == 1
        a.residual, a.res_ld, a.res_dtype = residual.data_ptr(), residual.stride(0), _lib.dtype_code(residual.dtype)
    a.act = int(act)
    if image_map is not None:
        a.map_mode = 1
        a.Hm, a.Wm, a.i0, a.j0, a.Hout, a.Wout = (int(v) for v in image_map)
    _lib.check(_lib.lib().pf_gemm_taps(C.byref(a), _st()))
    return out


def fmha(q: Tensor, k: Tensor, v: Tensor, out: Tensor, *, heads: int, head_dim: int, scale: float,
         bias: Optional[Tensor] = None) -> Tensor:
    """out[b, l, h*d:(h+1)*d] = softmax(q_h k_h^T * scale + bias) v_h; see pf_fmha_fwd.

    q: [B, Lq, >=H*d] view (last stride 1), k/v: [B, Lk, >=H*d] views — slices of a fused QKV buffer are fine.
    bias: fp32 [Lq, Lk] (shared by batches and heads) or [B, Lq, Lk].
    """
    _lib.require_cuda(q, k, v, out)
    a = FmhaArgs()
    B, Lq = q.shape[0], q.shape[1]
    Lk = k.shape[1]
    for t in (q, k, v, out):
        assert t.dim() == 3 and t.stride(2) == 1
    a.q, a.k, a.v, a.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.dtype = _lib.dtype_code(q.dtype)
    a.B, a.H, a.Lq, a.Lk, a.head_dim = B, int(heads), Lq, Lk, int(head_dim)
    a.q_ld, a.k_ld, a.v_ld, a.out_ld = q.stride(1), k.stride(1), v.stride(1), out.stride(1)
    a.q_bstride, a.k_bstride, a.v_bstride = q.stride(0), k.stride(0), v.stride(0)
    assert out.stride(0) == Lq * out.stride(1)
    a.scale = float(scale)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.stride(-1) == 1
        a.bias = bias.data_ptr()
        if bias.dim() == 3:
            a.bias_bstride, a.bias_ld = (bias.stride(0) if bias.shape[0] > 1 else 0), bias.stride(1)
        else:
            a.bias_bstride, a.bias_ld = 0, bias.stride(0)
    _lib.check(_lib.lib().pf_fmha_fwd(C.byref(a), _st()))
    return out
