"""Oracle: duck-typed Stable-Diffusion-2 `UNet2DConditionModel` in plain PyTorch fp32.

[3P restatement, parity unpinned] diffusers 0.24.0 is not installed and its source is not under
/root/reference (environment.yaml:13); this restates the published block semantics for the
`stabilityai/stable-diffusion-2-base` config. The reference consumes the UNet by ATTRIBUTE
WALK (models/pano/MVGenModel.py:52-295), so what matters is the attribute tree and the call conventions
`resnets[j](x, temb)`, `attentions[j](x, encoder_hidden_states=...).sample`, `downsamplers[j](x)`,
`upsamplers[j](x)`; parameter names equal diffusers' so real checkpoints map 1:1. Test infrastructure only.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

SD2_CONFIG = dict(
    in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
    attention_heads=(5, 10, 20, 20), cross_attention_dim=1024, norm_num_groups=32,
    down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
)

# a narrow config with the same topology for CPU-sized tests (channels stay multiples of 64 for the tap-GEMM)
TINY_CONFIG = dict(SD2_CONFIG, block_out_channels=(64, 128, 128, 128), attention_heads=(1, 2, 2, 2),
                   cross_attention_dim=64)


class Timesteps(nn.Module):
    """diffusers Timesteps(num_channels, flip_sin_to_cos=True, downscale_freq_shift=0)."""

    def __init__(self, num_channels):
        super().__init__()
        self.num_channels = num_channels

    def forward(self, timesteps):
        half = self.num_channels // 2
        exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half
        emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
        return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample):
        return self.linear_2(self.act(self.linear_1(sample)))


class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, groups=32, eps=1e-5):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x, temb):
        h = self.conv1(self.nonlinearity(self.norm1(x)))
        h = h + self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    """diffusers Attention with the plain AttnProcessor (what the LoRA processor becomes at inference)."""

    def __init__(self, query_dim, heads, dim_head, cross_attention_dim=None):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        kv = cross_attention_dim or query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(kv, inner, bias=False)
        self.to_v = nn.Linear(kv, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, x, encoder_hidden_states=None):
        ctx = x if encoder_hidden_states is None else encoder_hidden_states
        b, n, _ = x.shape
        split = lambda t: t.reshape(b, t.shape[1], self.heads, -1).permute(0, 2, 1, 3)
        q, k, v = split(self.to_q(x)), split(self.to_k(ctx)), split(self.to_v(ctx))
        s = torch.matmul(q, k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)
        o = torch.matmul(torch.softmax(s, dim=-1), v).permute(0, 2, 1, 3).reshape(b, n, -1)
        return self.to_out[1](self.to_out[0](o))


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, heads, dim_head, cross_attention_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, encoder_hidden_states):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), encoder_hidden_states) + x
        return self.ff(self.norm3(x)) + x


class Transformer2DModel(nn.Module):
    """use_linear_projection=True, one layer, GroupNorm(32, eps 1e-6)."""

    def __init__(self, heads, dim_head, in_channels, cross_attention_dim, groups=32):
        super().__init__()
        inner = heads * dim_head
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim)])
        self.proj_out = nn.Linear(inner, in_channels)

    def forward(self, hidden_states, encoder_hidden_states=None):
        b, c, h, w = hidden_states.shape
        x = self.norm(hidden_states).permute(0, 2, 3, 1).reshape(b, h * w, c)
        x = self.proj_in(x)
        for blk in self.transformer_blocks:
            x = blk(x, encoder_hidden_states)
        x = self.proj_out(x).reshape(b, h, w, c).permute(0, 3, 1, 2)
        return SimpleNamespace(sample=x + hidden_states)


class Downsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.channels = self.out_channels = channels
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.channels = self.out_channels = channels
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cin, cout, temb, layers, heads, ctx_dim, cross_attn, add_down, groups):
        super().__init__()
        self.has_cross_attention = cross_attn
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb, groups) for i in range(layers)])
        if cross_attn:
            self.attentions = nn.ModuleList(
                [Transformer2DModel(heads, cout // heads, cout, ctx_dim, groups) for _ in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None


class MidBlock(nn.Module):
    def __init__(self, c, temb, heads, ctx_dim, groups):
        super().__init__()
        self.has_cross_attention = True
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb, groups), ResnetBlock2D(c, c, temb, groups)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, c // heads, c, ctx_dim, groups)])


class UpBlock(nn.Module):
    def __init__(self, cin, cout, prev, temb, layers, heads, ctx_dim, cross_attn, add_up, groups):
        super().__init__()
        self.has_cross_attention = cross_attn
        res = []
        for i in range(layers):
            skip = cin if i == layers - 1 else cout
            rin = prev if i == 0 else cout
            res.append(ResnetBlock2D(rin + skip, cout, temb, groups))
        self.resnets = nn.ModuleList(res)
        if cross_attn:
            self.attentions = nn.ModuleList(
                [Transformer2DModel(heads, cout // heads, cout, ctx_dim, groups) for _ in range(layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None


class UNet2DConditionModel(nn.Module):
    def __init__(self, in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                 attention_heads=(5, 10, 20, 20), cross_attention_dim=1024, norm_num_groups=32,
                 down_block_types=SD2_CONFIG["down_block_types"], up_block_types=SD2_CONFIG["up_block_types"]):
        super().__init__()
        boc, g = tuple(block_out_channels), norm_num_groups
        temb = boc[0] * 4
        self.config = SimpleNamespace(in_channels=in_channels, out_channels=out_channels, block_out_channels=boc,
                                      layers_per_block=layers_per_block, attention_head_dim=tuple(attention_heads),
                                      cross_attention_dim=cross_attention_dim, norm_num_groups=g)
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_proj = Timesteps(boc[0])
        self.time_embedding = TimestepEmbedding(boc[0], temb)
        self.down_blocks = nn.ModuleList()
        out = boc[0]
        for i, typ in enumerate(down_block_types):
            cin, out = out, boc[i]
            self.down_blocks.append(DownBlock(cin, out, temb, layers_per_block, attention_heads[i], cross_attention_dim,
                                              typ.startswith("CrossAttn"), i != len(boc) - 1, g))
        self.mid_block = MidBlock(boc[-1], temb, attention_heads[-1], cross_attention_dim, g)
        self.up_blocks = nn.ModuleList()
        rev, rev_heads = boc[::-1], tuple(attention_heads)[::-1]
        out = rev[0]
        for i, typ in enumerate(up_block_types):
            prev, out = out, rev[i]
            cin = rev[min(i + 1, len(boc) - 1)]
            self.up_blocks.append(UpBlock(cin, out, prev, temb, layers_per_block + 1, rev_heads[i],
                                          cross_attention_dim, typ.startswith("CrossAttn"), i != len(boc) - 1, g))
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-5)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype


def build_unet(config: dict = SD2_CONFIG, seed: int = 0) -> UNet2DConditionModel:
    """Seeded synthetic weights (no SD-2 checkpoint offline): PyTorch default init under `seed`."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    net = UNet2DConditionModel(**config).eval()
    torch.random.set_rng_state(g)
    return net
