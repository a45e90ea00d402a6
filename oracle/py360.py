"""Oracle: `py360convert.e2p` — the pixel-space equirect -> perspective convention of the dataset path
(external/py360convert/e2p.py:6-43, utils.py:104-132,231-243; used by utils/pano.py:160-161 `Equirectangular.to_perspective`
and dataset/PanoDataset.py:138). Different from the tensor `e2p` of external/Perspective_and_Equirectangular: half-pixel
centres (`uv2coor`), x = right / y = up / z = forward with `u = -yaw`, longitude wrap-around, pole rows padded with the
first / last row rolled by W/2, and scipy's legacy 'wrap' boundary (period n - 1), which is restated here and pinned
against scipy.ndimage.map_coordinates itself in tests/test_oracle_golden.py. Test infrastructure only."""
from __future__ import annotations

import numpy as np


def rotation_matrix(rad, ax):
    """utils.py:231-243."""
    ax = np.array(ax, dtype=np.float64)
    ax = ax / np.sqrt((ax ** 2).sum())
    R = np.diag([np.cos(rad)] * 3)
    R = R + np.outer(ax, ax) * (1.0 - np.cos(rad))
    ax = ax * np.sin(rad)
    return R + np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])


def rotations(u_deg, v_deg, in_rot_deg=0.0):
    """(Rx, Ry, Ri) of e2p.py:28-30 + utils.py:104-114."""
    u, v, in_rot = -u_deg * np.pi / 180, v_deg * np.pi / 180, in_rot_deg * np.pi / 180
    Rx = rotation_matrix(v, [1, 0, 0])
    Ry = rotation_matrix(u, [0, 1, 0])
    Ri = rotation_matrix(in_rot, np.array([0, 0, 1.0]).dot(Rx).dot(Ry))
    return Rx, Ry, Ri


def coords(fov_deg, u_deg, v_deg, out_hw, h, w, in_rot_deg=0.0):
    """Sampling coordinates (coor_x, coor_y) in the equirect image: xyzpers -> xyz2uv -> uv2coor (utils.py:104-132)."""
    h_fov, v_fov = fov_deg[0] * np.pi / 180, fov_deg[1] * np.pi / 180
    out = np.ones((*out_hw, 3), np.float32)
    x_max, y_max = np.tan(h_fov / 2), np.tan(v_fov / 2)
    x_rng = np.linspace(-x_max, x_max, num=out_hw[1], dtype=np.float32)
    y_rng = np.linspace(-y_max, y_max, num=out_hw[0], dtype=np.float32)
    out[..., :2] = np.stack(np.meshgrid(x_rng, -y_rng), -1)
    Rx, Ry, Ri = rotations(u_deg, v_deg, in_rot_deg)
    xyz = out.dot(Rx).dot(Ry).dot(Ri)
    x, y, z = xyz[..., 0], xyz[..., 1], xyz[..., 2]
    u = np.arctan2(x, z)
    v = np.arctan2(y, np.sqrt(x ** 2 + z ** 2))
    return (u / (2 * np.pi) + 0.5) * w - 0.5, (-v / np.pi + 0.5) * h - 0.5


def _wrap(c, n):
    """scipy.ndimage 'wrap' (the legacy mode whose period is n - 1: first and last sample overlap)."""
    c = np.asarray(c, dtype=np.float64).copy()
    sz = n - 1
    neg = c < 0
    c[neg] += sz * ((-c[neg] / sz).astype(np.int64) + 1)
    big = c > n - 1
    c[big] -= sz * (c[big] / sz).astype(np.int64)
    return c


def sample_equirec(e_img, coor_x, coor_y, order):
    """utils.py:123-130: pole rows appended (last row, then first row, each rolled by W/2), then bilinear (order 1) or
    nearest (order 0) interpolation with 'wrap' boundaries; integer images are rounded half up like scipy does."""
    w = e_img.shape[1]
    pad_u = np.roll(e_img[[0]], w // 2, 1)
    pad_d = np.roll(e_img[[-1]], w // 2, 1)
    img = np.concatenate([e_img, pad_d, pad_u], 0)
    H, W = img.shape
    y, x = _wrap(coor_y, H), _wrap(coor_x, W)
    nb = lambda i, n: np.where(i > n - 1, i - (n - 1), i)
    if order == 0:
        val = img[nb(np.floor(y + 0.5).astype(np.int64), H), nb(np.floor(x + 0.5).astype(np.int64), W)]
        return val
    y0, x0 = np.floor(y).astype(np.int64), np.floor(x).astype(np.int64)
    ty, tx = y - y0, x - x0
    y1, x1 = nb(y0 + 1, H), nb(x0 + 1, W)
    f = img.astype(np.float64)
    val = f[y0, x0] * (1 - ty) * (1 - tx) + f[y0, x1] * (1 - ty) * tx + f[y1, x0] * ty * (1 - tx) + f[y1, x1] * ty * tx
    if np.issubdtype(e_img.dtype, np.integer):
        info = np.iinfo(e_img.dtype)
        return np.clip(np.floor(val + 0.5), info.min, info.max).astype(e_img.dtype)
    return val.astype(e_img.dtype)


def e2p(e_img, fov_deg, u_deg, v_deg, out_hw, in_rot_deg=0, mode="bilinear"):
    """e2p.py:6-43. e_img [H, W] or [H, W, C] numpy."""
    assert e_img.ndim in (2, 3)
    if mode not in ("bilinear", "nearest"):
        raise NotImplementedError("unknown mode")
    order = 1 if mode == "bilinear" else 0
    h, w = e_img.shape[:2]
    cx, cy = coords(fov_deg, u_deg, v_deg, out_hw, h, w, in_rot_deg)
    if e_img.ndim == 2:
        return sample_equirec(e_img, cx, cy, order)
    return np.stack([sample_equirec(e_img[..., i], cx, cy, order) for i in range(e_img.shape[2])], axis=-1)
