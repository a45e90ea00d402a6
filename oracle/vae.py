"""Oracle: the image-space ends of the loop — VAE decode with circular latent padding + `tensor_to_image` after sampling,
and the training step's `encode_image` with circular image padding (PanoGenerator.py:214-225, PanFusion.py:66-71) before it.

First-party code restated: `decode_latent` (models/pano/PanoGenerator.py:272-278),
`pad_pano(latent=True)` / `unpad_pano` around the panorama decode (PanFusion.py:166-172, PanoGenerator.py:227-238,
`latent_pad = 8`), `tensor_to_image` (models/modules/utils.py:9-15).

[3P restatement, parity unpinned] the decoder itself is diffusers 0.24.0 `AutoencoderKL` (not installed, source not
under /root/reference) with the `stabilityai/stable-diffusion-2-base` VAE config: latent_channels 4,
block_out_channels (128, 256, 512, 512), layers_per_block 2 (decoder: 3 resnets per up block), norm_num_groups 32,
resnet eps 1e-6, no time embedding, mid block = resnet / single-head attention (GroupNorm, q/k/v/out with bias,
residual) / resnet, scaling_factor 0.18215. Attribute / state-dict names equal diffusers'. Test infrastructure only.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from .eppa import pad_pano, unpad_pano

SD2_VAE_CONFIG = dict(latent_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                      norm_num_groups=32, scaling_factor=0.18215)
# narrow decoder with the same topology for CPU-sized tests (channels stay multiples of 64 for the tap-GEMM)
TINY_VAE_CONFIG = dict(SD2_VAE_CONFIG, block_out_channels=(64, 64, 128, 128))


class ResnetBlock2D(nn.Module):
    """diffusers ResnetBlock2D with temb_channels=None, eps 1e-6, output_scale_factor 1."""

    def __init__(self, cin, cout, groups, eps=1e-6):
        super().__init__()
        self.in_channels, self.out_channels = cin, cout
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb=None):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    """The VAE mid-block attention: one head of width C, GroupNorm on the input, residual connection."""

    def __init__(self, c, groups, eps=1e-6):
        super().__init__()
        self.heads = 1
        self.group_norm = nn.GroupNorm(groups, c, eps=eps)
        self.to_q, self.to_k, self.to_v = nn.Linear(c, c), nn.Linear(c, c), nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x.reshape(b, c, h * w)).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        p = torch.softmax(q @ k.transpose(1, 2) * (c ** -0.5), dim=-1)
        o = self.to_out[0](p @ v)
        return o.transpose(1, 2).reshape(b, c, h, w) + x


class Upsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.channels = self.out_channels = channels
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class UpDecoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, add_upsample, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_upsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                x = u(x)
        return x


class MidBlock2D(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, groups), ResnetBlock2D(c, c, groups)])
        self.attentions = nn.ModuleList([Attention(c, groups)])

    def forward(self, x):
        x = self.resnets[0](x)
        x = self.attentions[0](x)
        return self.resnets[1](x)


class Downsample2D(nn.Module):
    """diffusers Downsample2D(padding=0) as the VAE encoder uses it: zero-pad right and bottom by one, 3x3 stride-2 conv."""

    def __init__(self, channels):
        super().__init__()
        self.channels = self.out_channels = channels
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class DownEncoderBlock2D(nn.Module):
    def __init__(self, cin, cout, layers, add_downsample, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_downsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                x = d(x)
        return x


class Encoder(nn.Module):
    """diffusers Encoder (double_z): conv_in, DownEncoderBlock2D x len(block_out_channels), mid block, GroupNorm / SiLU /
    conv_out to 2 * latent_channels."""

    def __init__(self, in_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups):
        super().__init__()
        boc, g = tuple(block_out_channels), norm_num_groups
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        out = boc[0]
        for i, c in enumerate(boc):
            prev, out = out, c
            self.down_blocks.append(DownEncoderBlock2D(prev, out, layers_per_block, i != len(boc) - 1, g))
        self.mid_block = MidBlock2D(boc[-1], g)
        self.conv_norm_out = nn.GroupNorm(g, boc[-1], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[-1], 2 * latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for blk in self.down_blocks:
            x = blk(x)
        x = self.mid_block(x)
        return self.conv_out(self.conv_act(self.conv_norm_out(x)))


class DiagonalGaussianDistribution:
    """diffusers DiagonalGaussianDistribution: moments = [mean | logvar] along the channels, logvar clamped to [-30, 20]."""

    def __init__(self, parameters):
        self.mean, logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None, noise=None):
        if noise is None:
            noise = torch.randn(self.mean.shape, generator=generator, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


class Decoder(nn.Module):
    def __init__(self, latent_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups):
        super().__init__()
        boc, g = tuple(block_out_channels), norm_num_groups
        self.conv_in = nn.Conv2d(latent_channels, boc[-1], 3, padding=1)
        self.mid_block = MidBlock2D(boc[-1], g)
        self.up_blocks = nn.ModuleList()
        rev = boc[::-1]
        out = rev[0]
        for i, c in enumerate(rev):
            prev, out = out, c
            self.up_blocks.append(UpDecoderBlock2D(prev, out, layers_per_block + 1, i != len(rev) - 1, g))
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for blk in self.up_blocks:
            x = blk(x)
        return self.conv_out(self.conv_act(self.conv_norm_out(x)))


class AutoencoderKL(nn.Module):
    """`decode(z).sample` and `encode(x).latent_dist` like diffusers (`post_quant_conv`, `decoder`, `quant_conv`, `encoder`)."""

    def __init__(self, latent_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 norm_num_groups=32, scaling_factor=0.18215):
        super().__init__()
        self.config = SimpleNamespace(latent_channels=latent_channels, out_channels=out_channels,
                                      block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                                      norm_num_groups=norm_num_groups, scaling_factor=scaling_factor)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)
        self.decoder = Decoder(latent_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups)
        # the encoder half is created AFTER the decoder so that the decoder's seeded weights (build_vae) do not move
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.encoder = Encoder(out_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups)

    @property
    def dtype(self):
        return self.post_quant_conv.weight.dtype

    def decode(self, z):
        return SimpleNamespace(sample=self.decoder(self.post_quant_conv(z)))

    def encode(self, x):
        return SimpleNamespace(latent_dist=DiagonalGaussianDistribution(self.quant_conv(self.encoder(x))))


def build_vae(config: dict = SD2_VAE_CONFIG, seed: int = 21) -> AutoencoderKL:
    st = torch.random.get_rng_state()
    torch.manual_seed(seed)
    vae = AutoencoderKL(**config).eval()
    torch.random.set_rng_state(st)
    return vae


# ---- first-party code around the decoder -------------------------------------------------------------------

def decode_latent(latents, vae):
    """PanoGenerator.py:272-278: [b, m, 4, h, w] -> [b, m, 3, 8h, 8w]."""
    b = latents.shape[0]
    z = (1 / vae.config.scaling_factor * latents).flatten(0, 1)
    image = vae.decode(z.to(vae.dtype)).sample
    return image.reshape(b, -1, *image.shape[1:])


def encode_image(x_input, vae, generator=None, noise=None):
    """PanoGenerator.py:214-225: [b, l, 3, H, W] in [-1, 1] -> sampled latents [b, l, 4, H/8, W/8] * scaling_factor."""
    b = x_input.shape[0]
    dist = vae.encode(x_input.to(vae.dtype).flatten(0, 1)).latent_dist
    z = dist.sample(generator=generator, noise=noise)
    z = z.reshape(b, -1, *z.shape[1:])
    return z * vae.config.scaling_factor


def encode_pano(pano, vae, latent_pad: int = 8, generator=None, noise=None):
    """PanFusion.py:69-71: pad the IMAGE circularly by 8 * latent_pad pixels, encode, crop latent_pad latent columns."""
    return unpad_pano(encode_image(pad_pano(pano, 8 * latent_pad), vae, generator, noise), latent_pad)


def decode_pano(pano_latent, vae, latent_pad: int = 8):
    """PanFusion.py:169-171: circular padding of the LATENT by latent_pad columns, decode, crop 8*latent_pad pixels."""
    return unpad_pano(decode_latent(pad_pano(pano_latent, latent_pad), vae), 8 * latent_pad)


def tensor_to_image(image):
    """models/modules/utils.py:9-15: [-1, 1] float [..., c, h, w] -> uint8 numpy [..., h, w, c]."""
    if image.dtype != torch.uint8:
        image = (image / 2 + 0.5).clamp(0, 1)
        image = (image * 255).round()
    image = image.cpu().numpy().astype("uint8")
    return image.transpose(*range(image.ndim - 3), -2, -1, -3)
