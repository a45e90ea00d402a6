"""`pad_pano` / `unpad_pano` behind the reference's signatures (utils/pano.py:74-105).

Inside the denoiser the circular padding never materialises (engine.py folds it into the GroupNorm statistics, the
conv-prep kernel and the tap-GEMM's row map); these are the stand-alone functions the reference also uses around the
VAE (models/pano/PanoGenerator.py:227-238) and that callers of the drop-in may import.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import Tensor

from . import _lib


def pad_pano(pano: Tensor, padding: int) -> Tensor:
    """Circular padding of the last axis of a [b, c, h, w] or [b, m, c, h, w] tensor by `padding` columns per side."""
    if padding <= 0:
        return pano
    if pano.ndim not in (4, 5):
        raise NotImplementedError("pano should be 4 or 5 dim")
    _lib.require_cuda(pano)
    x = pano.contiguous()
    W = x.shape[-1]
    out = torch.empty((*x.shape[:-1], W + 2 * padding), dtype=x.dtype, device=x.device)
    rows = x.numel() // W
    _lib.check(_lib.lib().pf_pad_pano(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), x.element_size(),
                                      C.c_longlong(rows), W, int(padding), C.c_void_p(_lib.stream_ptr())))
    return out


def unpad_pano(pano_pad: Tensor, padding: int) -> Tensor:
    """Crop `padding` columns per side (a view, as in the reference)."""
    if padding <= 0:
        return pano_pad
    return pano_pad[..., padding:-padding]
