"""Equirectangular <-> perspective resampling behind the reference's `e2p` / `p2e` signatures.

Reference: external/Perspective_and_Equirectangular/e2p.py:54-76, p2e.py:52-77 (tensor path), utils.py:5-23.
The reference builds one float64 sampling grid per camera on the CPU and uploads it; here the host only
prepares a 20-double camera record (two 3x3 rotations + the two frustum half-extents) and the CUDA kernel
evaluates the grid on the fly (csrc/resample.cu).
"""
from __future__ import annotations

import functools
import math

import numpy as np
import torch
from torch import Tensor

from . import _lib


def _rodrigues(rvec: np.ndarray) -> np.ndarray:
    """Rotation vector -> matrix (what cv2.Rodrigues returns for a 3-vector), float64."""
    rvec = np.asarray(rvec, dtype=np.float64).reshape(3)
    theta = math.sqrt(float(rvec[0] * rvec[0] + rvec[1] * rvec[1] + rvec[2] * rvec[2]))
    if theta < np.finfo(np.float64).eps:
        return np.eye(3)
    c, s = math.cos(theta), math.sin(theta)
    k = rvec / theta
    kx = np.array([[0.0, -k[2], k[1]], [k[2], 0.0, -k[0]], [-k[1], k[0], 0.0]])
    return c * np.eye(3) + (1.0 - c) * np.outer(k, k) + s * kx


@functools.lru_cache(maxsize=4096)
def _camera_record(kind: str, fov: float, theta: float, phi: float, h: int, w: int) -> tuple:
    """20 doubles: R1, R2 (row-major), w_len, h_len.  kind 'e2p': forward rotations (e2p.py:23-30);
    kind 'p2e': their inverses (p2e.py:23-29). (h, w) is the perspective image size (hfov = h/w * wfov)."""
    hfov = float(h) / w * fov
    w_len = np.tan(np.radians(fov / 2.0))
    h_len = np.tan(np.radians(hfov / 2.0))
    y_axis = np.array([0.0, 1.0, 0.0])
    z_axis = np.array([0.0, 0.0, 1.0])
    R1 = _rodrigues(z_axis * np.radians(theta))
    R2 = _rodrigues(np.dot(R1, y_axis) * np.radians(-phi))
    if kind == "p2e":
        R1 = np.linalg.inv(R1)
        R2 = np.linalg.inv(R2)
    rec = np.concatenate([R1.reshape(-1), R2.reshape(-1), [w_len, h_len]]).astype(np.float64)
    return tuple(rec.tolist())


_RECORD_CACHE: dict = {}


def _scalar(v, i):
    """index_list_or_scalar (utils.py:18-23)."""
    if hasattr(v, "__len__"):
        v = v[i]
    if isinstance(v, Tensor):
        v = v.item()
    return float(v)


def _values(v, n: int) -> tuple:
    """index_list_or_scalar (utils.py:18-23) for i in range(n), with ONE host read per argument (a tensor argument
    costs one .tolist(), not n .item() calls — the launch path of a 70 us kernel must not take longer than the kernel)."""
    if isinstance(v, Tensor):
        v = v.detach().reshape(-1).tolist() if v.dim() else float(v)
    if hasattr(v, "__len__"):
        return tuple(_scalar(v, i) for i in range(n))
    return (float(v),) * n


def camera_records(kind: str, fov_deg, u_deg, v_deg, batch: int, h: int, w: int, device) -> tuple[Tensor, int]:
    """-> (records[n, 20] float64 on device, cam_stride). All-scalar cameras broadcast (e2p.py:65-66)."""
    if all(not hasattr(v, "__len__") for v in (fov_deg, u_deg, v_deg)):
        n, stride = 1, 0
    else:
        n, stride = batch, 1
    key = (kind, _values(fov_deg, n), _values(u_deg, n), _values(v_deg, n), int(h), int(w), str(device))
    t = _RECORD_CACHE.get(key)
    if t is None:
        recs = [_camera_record(kind, key[1][i], key[2][i], key[3][i], int(h), int(w)) for i in range(n)]
        t = torch.tensor(recs, dtype=torch.float64).to(device)
        if len(_RECORD_CACHE) > 4096:
            _RECORD_CACHE.clear()
        _RECORD_CACHE[key] = t  # device-resident: repeated warps with the same rig launch without any host->device copy
    return t, stride


def _mode_code(mode) -> int:
    # choose_mode (utils.py:5-8): tensors default to bilinear
    mode = mode if mode else "bilinear"
    if mode == "bilinear":
        return 0
    if mode == "nearest":
        return 1
    raise ValueError("mode must be one of [bilinear, bicubic, nearest]")


def e2p(e_img: Tensor, fov_deg, u_deg, v_deg, out_hw, mode=None, views_per_image: int = 1) -> Tensor:
    """Equirect [B,C,He,We] -> perspective [B,C,h,w] (e2p.py:54-76).

    views_per_image > 1 (extension): e_img is [B / views_per_image, C, He, We] and camera b looks at image
    b // views_per_image — what the reference gets by expanding one panorama to its m views before the call
    (PanFusion.py:33-37), without materialising the copies."""
    _lib.require_cuda(e_img)
    bs, c, he, we = e_img.shape
    b = bs * int(views_per_image)
    h, w = int(out_hw[0]), int(out_hw[1])
    e_img = e_img.contiguous()
    cams, stride = camera_records("e2p", fov_deg, u_deg, v_deg, b, h, w, e_img.device)
    out = torch.empty((b, c, h, w), dtype=e_img.dtype, device=e_img.device)
    args = (_lib.C.c_void_p(e_img.data_ptr()), _lib.C.c_void_p(out.data_ptr()), _lib.dtype_code(e_img.dtype))
    tail = (c, he, we, h, w, _lib.C.c_void_p(cams.data_ptr()), stride, _mode_code(mode), _lib.C.c_void_p(_lib.stream_ptr()))
    if views_per_image == 1:
        _lib.check(_lib.lib().pf_e2p(*args, b, *tail))
    else:
        if not stride:
            raise ValueError("views_per_image needs per-view cameras")
        _lib.check(_lib.lib().pf_e2p_shared(*args, b, int(views_per_image), *tail))
    return out


def p2e(p_img: Tensor, fov_deg, u_deg, v_deg, out_hw, mode=None):
    """Perspective [B,C,hp,wp] -> (equirect [B,C,H,W] * mask, mask [B,1,H,W] bool) (p2e.py:52-77)."""
    _lib.require_cuda(p_img)
    b, c, hp, wp = p_img.shape
    H, W = int(out_hw[0]), int(out_hw[1])
    p_img = p_img.contiguous()
    cams, stride = camera_records("p2e", fov_deg, u_deg, v_deg, b, hp, wp, p_img.device)
    out = torch.empty((b, c, H, W), dtype=p_img.dtype, device=p_img.device)
    nmask = b if stride else 1
    mask = torch.empty((b, 1, H, W), dtype=torch.uint8, device=p_img.device)
    _lib.check(_lib.lib().pf_p2e(
        _lib.C.c_void_p(p_img.data_ptr()), _lib.C.c_void_p(out.data_ptr()), _lib.C.c_void_p(mask.data_ptr()),
        _lib.dtype_code(p_img.dtype), b, c, hp, wp, H, W, _lib.C.c_void_p(cams.data_ptr()), stride,
        _mode_code(mode), _lib.C.c_void_p(_lib.stream_ptr())))
    mask = mask.bool()
    if not stride:
        mask = mask[:nmask]  # reference returns a [1,1,H,W] mask when cameras are scalars
    return out, mask
