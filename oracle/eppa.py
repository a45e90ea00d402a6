"""Oracle: EPPA (Equirectangular-Perspective Projection Attention) — plain PyTorch fp32 restatement.

Follows models/pano/utils.py:10-106 (get_masks, get_coords), models/modules/transformer.py:8-74,130-201
(GEGLU, FeedForward, CrossAttention, BasicTransformerBlock, SphericalPE) and models/pano/modules.py:8-59
(WarpAttn). Third-party pieces restated from published behaviour: kornia gaussian_blur2d / create_meshgrid,
xformers memory_efficient_attention ([3P], see oracle/__init__.py). Module / parameter names equal the
reference's so state dicts are interchangeable. Test infrastructure only.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import geometry as G


# ------------------------------------------------------------------------------------------------
# circular padding (utils/pano.py:74-105)
# ------------------------------------------------------------------------------------------------
def pad_pano(pano: Tensor, padding: int) -> Tensor:
    if padding <= 0:
        return pano
    if pano.ndim not in (4, 5):
        raise NotImplementedError("pano should be 4 or 5 dim")
    lead = pano.shape[:-2]
    flat = pano.reshape(-1, *pano.shape[-2:])
    flat = F.pad(flat, [padding, padding], mode="circular")
    return flat.reshape(*lead, *flat.shape[-2:])


def unpad_pano(pano_pad: Tensor, padding: int) -> Tensor:
    if padding <= 0:
        return pano_pad
    return pano_pad[..., padding:-padding]


# ------------------------------------------------------------------------------------------------
# kornia.filters.gaussian_blur2d((5,5),(1,1), border_type='replicate', separable=True) [3P]
# ------------------------------------------------------------------------------------------------
def gaussian_kernel1d(ksize: int = 5, sigma: float = 1.0, dtype=torch.float32) -> Tensor:
    x = torch.arange(ksize, dtype=dtype) - ksize // 2
    g = torch.exp(-x.pow(2.0) / (2 * sigma ** 2))
    return g / g.sum()


def gaussian_blur5_replicate(x: Tensor) -> Tensor:
    """x: [n,1,h,w]. Horizontal pass then vertical pass, each with replicate padding of 2."""
    k = gaussian_kernel1d(5, 1.0, x.dtype).to(x.device)
    x = F.conv2d(F.pad(x, [2, 2, 0, 0], mode="replicate"), k.view(1, 1, 1, 5))
    x = F.conv2d(F.pad(x, [0, 0, 2, 2], mode="replicate"), k.view(1, 1, 5, 1))
    return x


# ------------------------------------------------------------------------------------------------
# get_masks (models/pano/utils.py:10-84)
# ------------------------------------------------------------------------------------------------
def get_masks(pers_h, pers_w, equi_h, equi_w, cameras, device="cpu", dtype=torch.float32):
    """-> pers_masks [m, eh, ew, ph, pw], equi_masks [m, ph, pw, eh, ew], values in [-1, 1]."""
    m = len(cameras["FoV"])
    P, E = pers_h * pers_w, equi_h * equi_w
    # one image per source pixel, holding a single 1 at that pixel (utils.py:20-29)
    pers_onehot = torch.eye(P, dtype=dtype).reshape(1, P, pers_h, pers_w).repeat(m, 1, 1, 1)
    equi_onehot = torch.eye(E, dtype=dtype).reshape(1, E, equi_h, equi_w).repeat(m, 1, 1, 1)
    fov, theta, phi = cameras["FoV"], cameras["theta"], cameras["phi"]
    # warp the indicator stacks (utils.py:34-41)
    equi_masks = G.p2e(pers_onehot, fov, theta, phi, (equi_h, equi_w))[0]       # [m, P, eh, ew]
    pers_masks = G.e2p(equi_onehot, fov, theta, phi, (pers_h, pers_w))          # [m, E, ph, pw]
    # symmetric union: what either warp says corresponds (utils.py:52-60); second update sees the first
    pm = pers_masks.reshape(m, E, P)
    em = equi_masks.reshape(m, P, E)
    pm = (pm + em.transpose(1, 2)).clamp(0, 1)
    em = (em + pm.transpose(1, 2)).clamp(0, 1)
    # blur over the key image (utils.py:63-68): perspective keys replicate border; equirect keys wrap in W
    pm = gaussian_blur5_replicate(pm.reshape(m * E, 1, pers_h, pers_w))
    em = em.reshape(m * P, 1, equi_h, equi_w)
    em = unpad_pano(gaussian_blur5_replicate(pad_pano(em, 2)), 2)
    # per-query normalisation to [-1, 1] (utils.py:69-76)
    def _norm(t):
        mx = torch.amax(t, dim=(1, 2, 3), keepdim=True)
        mx[mx == 0] = 1.0
        return t / mx * 2 - 1
    pm, em = _norm(pm), _norm(em)
    return (pm.reshape(m, equi_h, equi_w, pers_h, pers_w).to(device),
            em.reshape(m, pers_h, pers_w, equi_h, equi_w).to(device))


# ------------------------------------------------------------------------------------------------
# get_coords (models/pano/utils.py:87-106)
# ------------------------------------------------------------------------------------------------
def get_coords(pers_h, pers_w, equi_h, equi_w, cameras, device="cpu", dtype=torch.float32):
    """-> pers_coords [m, ph, pw, 2] (lon, lat down-positive), equi_coords [eh, ew, 2] (lon, lat up-positive)."""
    lon, lat = np.meshgrid(np.linspace(-np.pi, np.pi, equi_w), np.linspace(np.pi / 2, -np.pi / 2, equi_h))
    equi_coords = torch.tensor(np.stack([lon, lat], -1), device=device, dtype=dtype)
    pers = []
    for fov, theta, phi in zip(cameras["FoV"], cameras["theta"], cameras["phi"]):
        plon, plat = G.pers_coords_to_equi(fov.item(), theta.item(), phi.item(), pers_h, pers_w)
        pers.append(torch.tensor(np.stack([plon, plat], -1), device=device, dtype=dtype))
    return torch.stack(pers, 0), equi_coords


# ------------------------------------------------------------------------------------------------
# transformer pieces (models/modules/transformer.py)
# ------------------------------------------------------------------------------------------------
class GEGLU(nn.Module):  # transformer.py:8-16
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        v, gate = self.proj(x).chunk(2, dim=-1)
        return v * F.gelu(gate)


class FeedForward(nn.Module):  # transformer.py:18-38 (glu=True branch is the one WarpAttn uses)
    def __init__(self, dim, mult=4):
        super().__init__()
        inner = int(dim * mult)
        proj = GEGLU(dim, inner)  # created first so seeded initialisation matches the reference's draw order
        out = nn.Linear(inner, dim)
        nn.init.zeros_(out.weight)
        nn.init.zeros_(out.bias)
        self.net = nn.Sequential(proj, nn.Dropout(0.0), out)

    def forward(self, x):
        return self.net(x)


class CrossAttention(nn.Module):  # transformer.py:41-74
    def __init__(self, query_dim, context_dim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Linear(inner, query_dim)
        nn.init.zeros_(self.to_out.weight)
        nn.init.zeros_(self.to_out.bias)

    def forward(self, x, context, mask):
        b, n, _ = x.shape
        h = self.heads
        split = lambda t: t.reshape(b, t.shape[1], h, -1).permute(0, 2, 1, 3)  # b h n d
        q, k, v = split(self.to_q(x)), split(self.to_k(context)), split(self.to_v(context))
        # xformers memory_efficient_attention(q, k, v, attn_bias) [3P]: softmax(q k^T / sqrt(d) + bias) v,
        # the bias shared by all heads (transformer.py:68)
        s = torch.matmul(q, k.transpose(-1, -2)) * (q.shape[-1] ** -0.5) + mask[:, None]
        o = torch.matmul(torch.softmax(s, dim=-1), v)
        return self.to_out(o.permute(0, 2, 1, 3).reshape(b, n, -1))


class BasicTransformerBlock(nn.Module):  # transformer.py:130-162 (forward == _forward; checkpointing is a no-op here)
    def __init__(self, dim, n_heads, d_head, context_dim):
        super().__init__()
        self.attn1 = CrossAttention(dim, context_dim, n_heads, d_head)
        self.ff = FeedForward(dim)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)

    def forward(self, x, context, mask, query_pe):
        q = self.norm1(x + query_pe)
        ctx = self.norm1(context)
        x = self.attn1(q, ctx, mask) + x
        return self.ff(self.norm2(x)) + x


class SphericalPE(nn.Module):  # transformer.py:165-201
    def __init__(self, n_freqs):
        super().__init__()
        base = 2 if n_freqs <= 80 else 5000 ** (1 / (n_freqs / 2.5))
        self.register_buffer("freq_bands", base ** torch.linspace(0, n_freqs - 1, n_freqs))

    def forward(self, coords):
        lead = coords.shape[:-1]
        ang = coords.reshape(-1, 2, 1) * self.freq_bands           # [n, 2, N]
        pe = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)     # [n, 4, N]: sin lon, sin lat, cos lon, cos lat
        return pe.reshape(*lead, -1)


class WarpAttn(nn.Module):  # models/pano/modules.py:8-59
    def __init__(self, dim):
        super().__init__()
        self.transformer = BasicTransformerBlock(dim, dim // 32, 32, context_dim=dim)
        self.pe = SphericalPE(dim // 4)

    def forward(self, pers_x, equi_x, cameras):
        bm, c, ph, pw = pers_x.shape
        b, _, eh, ew = equi_x.shape
        m = bm // b
        pers_masks, equi_masks = get_masks(ph, pw, eh, ew, cameras, pers_x.device, pers_x.dtype)
        pers_coords, equi_coords = get_coords(ph, pw, eh, ew, cameras, pers_x.device, pers_x.dtype)
        pers_pe = self.pe(pers_coords)                                   # [bm, ph, pw, c]
        equi_pe = self.pe(equi_coords)[None].expand(b, -1, -1, -1)       # [b, eh, ew, c]
        pers_tok = pers_x.permute(0, 2, 3, 1).reshape(b, m * ph * pw, c)
        equi_tok = equi_x.permute(0, 2, 3, 1).reshape(b, eh * ew, c)
        pers_pe_tok = pers_pe.reshape(b, m * ph * pw, c)
        equi_pe_tok = equi_pe.reshape(b, eh * ew, c)
        # perspective -> equirect (modules.py:44-48): queries = pano tokens, keys = all views' tokens
        bias1 = pers_masks.reshape(b, m, eh * ew, ph * pw).permute(0, 2, 1, 3).reshape(b, eh * ew, m * ph * pw)
        equi_out = self.transformer(equi_tok, pers_tok + pers_pe_tok, bias1, equi_pe_tok)
        # equirect -> perspective (modules.py:51-55); reads the INPUT features, not equi_out
        bias2 = equi_masks.reshape(b, m * ph * pw, eh * ew)
        pers_out = self.transformer(pers_tok, equi_tok + equi_pe_tok, bias2, pers_pe_tok)
        pers_out = pers_out.reshape(bm, ph, pw, c).permute(0, 3, 1, 2)
        equi_out = equi_out.reshape(b, eh, ew, c).permute(0, 3, 1, 2)
        return pers_out, equi_out
