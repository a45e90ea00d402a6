"""`py360convert.e2p` behind its own signature, on the GPU.

Reference: external/py360convert/e2p.py:6-43 (+ utils.py:104-132,231-243), called by `Equirectangular.to_perspective`
(utils/pano.py:160-161) once per view in the dataset (dataset/PanoDataset.py:138: 20 views per panorama on the CPU, scipy
map_coordinates per channel). Here one launch resamples ALL requested views of a panorama; only the three 3x3 rotations per
camera are prepared on the host (29 doubles, cached). numpy in -> numpy out like the reference; a CUDA tensor in -> CUDA
tensor out (no host round trip). There is no CPU fallback.
"""
from __future__ import annotations

import functools

import numpy as np
import torch

from . import _lib


def _rotation_matrix(rad: float, ax) -> np.ndarray:  # utils.py:231-243
    ax = np.array(ax, dtype=np.float64)
    ax = ax / np.sqrt((ax ** 2).sum())
    R = np.diag([np.cos(rad)] * 3)
    R = R + np.outer(ax, ax) * (1.0 - np.cos(rad))
    ax = ax * np.sin(rad)
    return R + np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])


@functools.lru_cache(maxsize=4096)
def _record(h_fov_deg: float, v_fov_deg: float, u_deg: float, v_deg: float, in_rot_deg: float) -> tuple:
    u, v, in_rot = -u_deg * np.pi / 180, v_deg * np.pi / 180, in_rot_deg * np.pi / 180  # e2p.py:28-29
    Rx = _rotation_matrix(v, [1, 0, 0])
    Ry = _rotation_matrix(u, [0, 1, 0])
    Ri = _rotation_matrix(in_rot, np.array([0, 0, 1.0]).dot(Rx).dot(Ry))
    ext = [np.tan(h_fov_deg * np.pi / 180 / 2), np.tan(v_fov_deg * np.pi / 180 / 2)]
    return tuple(np.concatenate([Rx.reshape(-1), Ry.reshape(-1), Ri.reshape(-1), ext]).tolist())


def _fov_pair(fov_deg):
    try:
        return float(fov_deg[0]), float(fov_deg[1])
    except TypeError:  # the reference's scalar branch is broken (NameError, e2p.py:17-18); a scalar means a square FoV
        return float(fov_deg), float(fov_deg)


def e2p_views(e_img, fov_deg, u_deg, v_deg, out_hw, in_rot_deg=0, mode="bilinear"):
    """All views of one panorama in one launch: u_deg / v_deg sequences of length m -> [m, h, w, C] ([m, h, w] for a 2-D image)."""
    if mode == "bilinear":
        code = 0
    elif mode == "nearest":
        code = 1
    else:
        raise NotImplementedError("unknown mode")
    as_numpy = isinstance(e_img, np.ndarray)
    x = torch.from_numpy(np.ascontiguousarray(e_img)).cuda() if as_numpy else e_img
    _lib.require_cuda(x)
    if x.dim() not in (2, 3):
        raise AssertionError("e_img must be [H, W] or [H, W, C]")
    squeeze = x.dim() == 2
    x = x.contiguous() if not squeeze else x.contiguous()[..., None]
    if x.dtype == torch.uint8:
        is_u8 = 1
    elif x.dtype == torch.float32:
        is_u8 = 0
    else:
        raise TypeError(f"py360 e2p: uint8 or float32 images, got {x.dtype}")
    H, W, C = x.shape
    h, w = int(out_hw[0]), int(out_hw[1])
    hf, vf = _fov_pair(fov_deg)
    us, vs = np.atleast_1d(np.asarray(u_deg, dtype=np.float64)), np.atleast_1d(np.asarray(v_deg, dtype=np.float64))
    rot = np.broadcast_to(np.asarray(in_rot_deg, dtype=np.float64), us.shape)
    recs = torch.tensor([_record(hf, vf, float(a), float(b), float(r)) for a, b, r in zip(us, vs, rot)],
                        dtype=torch.float64).to(x.device)
    out = torch.empty((len(us), h, w, C), dtype=x.dtype, device=x.device)
    Cv = _lib.C.c_void_p
    _lib.check(_lib.lib().pf_e2p_py360(Cv(x.data_ptr()), Cv(out.data_ptr()), is_u8, H, W, C, h, w, Cv(recs.data_ptr()),
                                       len(us), code, Cv(_lib.stream_ptr())))
    if squeeze:
        out = out[..., 0]
    return out.cpu().numpy() if as_numpy else out


def e2p(e_img, fov_deg, u_deg, v_deg, out_hw, in_rot_deg=0, mode="bilinear"):
    """py360convert.e2p(e_img[H, W, *], fov_deg, u_deg, v_deg, out_hw, in_rot_deg, mode) -> [h, w, *] (e2p.py:6-43)."""
    return e2p_views(e_img, fov_deg, [u_deg], [v_deg], out_hw, in_rot_deg, mode)[0]
