"""Host-side weight packing for the tap-GEMM (done once per model, not in the hot loop)."""
from __future__ import annotations

import torch
from torch import Tensor


def pack_conv3x3(w: Tensor) -> Tensor:
    """[Cout, Cin, kh, kw] -> [Cout, kh*kw*Cin] (tap-major, channel-minor): K-slab t*Cin.. matches tap t = ky*kw+kx."""
    co, ci, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(co, kh * kw * ci).contiguous()


def pack_geglu(w: Tensor, b: Tensor | None, block_n: int):
    """GEGLU projection [2*inner, K] (value rows then gate rows, transformer.py:14-15) -> rows interleaved so each
    block_n-row tile holds block_n/2 value rows followed by the matching block_n/2 gate rows."""
    n2, k = w.shape
    inner = n2 // 2
    half = block_n // 2
    assert inner % half == 0, (inner, half)
    val, gate = w[:inner], w[inner:]
    wp = torch.stack([val.reshape(inner // half, half, k), gate.reshape(inner // half, half, k)], dim=1)
    wp = wp.reshape(n2, k).contiguous()
    bp = None
    if b is not None:
        bv, bg = b[:inner], b[inner:]
        bp = torch.stack([bv.reshape(inner // half, half), bg.reshape(inner // half, half)], dim=1).reshape(n2)
        bp = bp.contiguous().float()
    return wp, bp


def pack_upsample_phases(w: Tensor) -> list[Tensor]:
    """Upsample2D = nearest x2 then a 3x3 convolution (pad 1). Output pixel (2i+a, 2j+b) only ever sees the 2x2 source
    neighbourhood rows {i-1+a, i+a} x cols {j-1+b, j+b}; the 3x3 taps that land on the same source pixel are summed (in
    fp32, before rounding to the compute type). -> four packed weights [Cout, 4*Cin], phase order (a, b) = (0,0), (0,1),
    (1,0), (1,1), tap order (r, c) row-major, matching engine.taps2x2."""
    w = w.detach().to(torch.float32)
    rows = {0: ([0], [1, 2]), 1: ([0, 1], [2])}  # phase a -> kernel rows feeding source row r = 0 / 1
    out = []
    for a in (0, 1):
        for b in (0, 1):
            taps = []
            for r in (0, 1):
                for c in (0, 1):
                    taps.append(w[:, :, rows[a][r]][:, :, :, rows[b][c]].sum(dim=(2, 3)))  # [Cout, Cin]
            out.append(torch.cat(taps, dim=1).contiguous())
    return out
