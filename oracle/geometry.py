"""Oracle: equirect <-> perspective resampling (numpy float64 grids + torch grid_sample).

Restates external/Perspective_and_Equirectangular/e2p.py:9-76 and p2e.py:9-77 (tensor path) together with the
kornia 0.7.2 `remap` they call ([3P]: normalize_pixel_coordinates + F.grid_sample(padding_mode='zeros',
align_corners=True)). Test infrastructure only.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor


def rodrigues(rvec) -> np.ndarray:
    """cv2.Rodrigues(3-vector)[0] in float64 (used at e2p.py:25-26, p2e.py:25-26)."""
    r = np.asarray(rvec, dtype=np.float64).reshape(3)
    theta = float(np.sqrt(r @ r))
    if theta < np.finfo(np.float64).eps:
        return np.eye(3)
    k = r / theta
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]], dtype=np.float64)
    return np.cos(theta) * np.eye(3) + (1 - np.cos(theta)) * np.outer(k, k) + np.sin(theta) * K


def camera_rotations(theta_deg: float, phi_deg: float):
    """R1 = yaw about z, R2 = pitch about the yawed y axis (e2p.py:23-26)."""
    z = np.array([0.0, 0.0, 1.0], np.float32)
    y = np.array([0.0, 1.0, 0.0], np.float32)
    R1 = rodrigues(z * np.radians(theta_deg))
    R2 = rodrigues(np.dot(R1, y) * np.radians(-phi_deg))
    return R1, R2


def pers_coords_to_equi(wfov, theta, phi, h, w):
    """(lon, lat) [h,w] float64 of every perspective pixel; lat is down-positive (e2p.py:9-36)."""
    hfov = float(h) / w * wfov
    w_len = np.tan(np.radians(wfov / 2.0))
    h_len = np.tan(np.radians(hfov / 2.0))
    ys = np.linspace(-w_len, w_len, w)[None, :].repeat(h, 0)
    zs = -np.linspace(-h_len, h_len, h)[:, None].repeat(w, 1)
    xs = np.ones((h, w), np.float32)
    norm = np.sqrt(xs ** 2 + ys ** 2 + zs ** 2)
    rays = np.stack([xs, ys, zs], -1) / norm[..., None]
    R1, R2 = camera_rotations(theta, phi)
    v = rays.reshape(-1, 3).T
    v = np.dot(R2, np.dot(R1, v)).T
    lat = np.arcsin(v[:, 2]).reshape(h, w)
    lon = np.arctan2(v[:, 1], v[:, 0]).reshape(h, w)
    return lon, -lat


def pers_pix_to_equi(eh, ew, fov, theta, phi, h, w):
    """Sampling positions (x, y) in equirect pixel units for every perspective pixel (e2p.py:39-51)."""
    lon, lat = pers_coords_to_equi(fov, theta, phi, h, w)
    cx, cy = (ew - 1) / 2.0, (eh - 1) / 2.0
    lon = lon / np.pi * 180
    lat = lat / np.pi * 180
    return lon / 180 * cx + cx, lat / 90 * cy + cy


def equi_pix_to_pers(ph, pw, wfov, theta, phi, h, w):
    """Sampling positions in perspective pixel units for every equirect pixel + validity (p2e.py:9-49)."""
    hfov = float(ph) / pw * wfov
    w_len = np.tan(np.radians(wfov / 2.0))
    h_len = np.tan(np.radians(hfov / 2.0))
    lon_d, lat_d = np.meshgrid(np.linspace(-180, 180, w), np.linspace(90, -90, h))
    lon_r, lat_r = np.radians(lon_d), np.radians(lat_d)
    d = np.stack([np.cos(lon_r) * np.cos(lat_r), np.sin(lon_r) * np.cos(lat_r), np.sin(lat_r)], -1)
    R1, R2 = camera_rotations(theta, phi)
    R1i, R2i = np.linalg.inv(R1), np.linalg.inv(R2)
    v = d.reshape(-1, 3).T
    v = np.dot(R1i, np.dot(R2i, v)).T.reshape(h, w, 3)
    front = v[..., 0] > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        u = v[..., 1] / v[..., 0]
        t = v[..., 2] / v[..., 0]
    inside = (-w_len < u) & (u < w_len) & (-h_len < t) & (t < h_len)
    with np.errstate(invalid="ignore"):
        xmap = np.where(inside, (u + w_len) / 2 / w_len * pw, 0)
        ymap = np.where(inside, (-t + h_len) / 2 / h_len * ph, 0)
    return xmap, ymap, inside & front


def remap(image: Tensor, map_x: Tensor, map_y: Tensor, mode: str = "bilinear") -> Tensor:
    """kornia.geometry.transform.remap(align_corners=True) [3P, kornia 0.7.2]."""
    b, _, h, w = image.shape
    grid = torch.stack([map_x, map_y], -1)  # (n, h', w', 2) in pixel units
    size = torch.tensor([w, h], dtype=grid.dtype, device=grid.device)
    factor = torch.tensor(2.0, dtype=grid.dtype) / (size - 1).clamp(1e-14)
    grid = factor * grid - 1
    return F.grid_sample(image, grid.expand(b, -1, -1, -1), mode=mode, padding_mode="zeros", align_corners=True)


def _item(v, i):
    if hasattr(v, "__len__"):
        v = v[i]
    if isinstance(v, Tensor):
        v = v.item()
    return v


def _all_scalar(*vs):
    return all(not hasattr(v, "__len__") for v in vs)


def e2p(e_img: Tensor, fov_deg, u_deg, v_deg, out_hw, mode=None) -> Tensor:
    """e2p.py:54-76, tensor path."""
    mode = mode or "bilinear"
    b, _, he, we = e_img.shape
    n = 1 if _all_scalar(fov_deg, u_deg, v_deg) else b
    xs, ys = zip(*[pers_pix_to_equi(he, we, _item(fov_deg, i), _item(u_deg, i), _item(v_deg, i), *out_hw)
                   for i in range(n)])
    mx = torch.from_numpy(np.stack(xs)).to(e_img.dtype)
    my = torch.from_numpy(np.stack(ys)).to(e_img.dtype)
    return remap(e_img, mx, my, mode)


def p2e(p_img: Tensor, fov_deg, u_deg, v_deg, out_hw, mode=None):
    """p2e.py:52-77, tensor path. Returns (equi * mask, mask[n,1,H,W] bool)."""
    mode = mode or "bilinear"
    b, _, hp, wp = p_img.shape
    n = 1 if _all_scalar(fov_deg, u_deg, v_deg) else b
    xs, ys, ms = zip(*[equi_pix_to_pers(hp, wp, _item(fov_deg, i), _item(u_deg, i), _item(v_deg, i), *out_hw)
                       for i in range(n)])
    mx = torch.from_numpy(np.stack(xs)).to(p_img.dtype)
    my = torch.from_numpy(np.stack(ys)).to(p_img.dtype)
    mask = torch.from_numpy(np.stack(ms)[:, None])
    return remap(p_img, mx, my, mode) * mask, mask
