"""Thin Python wrappers over the C-ABI kernels (raw pointers + current CUDA stream). No math happens here."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch
from torch import Tensor

from . import _lib
from ._lib import PF_ACT_GEGLU, PF_ACT_GELU, PF_ACT_NONE, PF_ACT_SILU, GemmArgs  # noqa: F401


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
