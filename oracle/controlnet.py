"""Oracle: duck-typed diffusers `ControlNetModel` in plain PyTorch fp32.

[3P restatement, parity unpinned] diffusers 0.24.0 is not installed and its source is not under
/root/reference (environment.yaml:13). The reference builds it with `ControlNetModel.from_unet(unet)`
(models/pano/PanoGenerator.py:153-157) and calls it as a black box
`cn(sample, timestep, encoder_hidden_states=..., controlnet_cond=..., return_dict=False)
 -> (down_block_res_samples[12], mid_block_res_sample)` (models/pano/MVGenModel.py:66-83); the first-party code that
consumes the result (residuals added to the SKIP tensors after the encoder and to the mid output,
MVGenModel.py:154-170,200-203) is pinned by executing the reference file (oracle/make_golden.py).

Published semantics restated here (diffusers 0.24.0 `models/controlnet.py`):
  * `ControlNetConditioningEmbedding(320, block_out_channels=(16, 32, 96, 256))`:
    conv_in 3->16, then per level [conv c->c, conv c->c' stride 2], SiLU after every conv but the last,
    zero-initialised conv_out 256->320 (all 3x3, padding 1);
  * encoder = copy of the UNet's conv_in / time embedding / down_blocks / mid_block, run with ZERO padding on the
    un-padded latent, `sample = conv_in(sample) + cond_embedding(cond)`;
  * one zero-initialised 1x1 conv per skip tensor (12) and one for the mid output; outputs scaled by
    `conditioning_scale` (1.0).
Test infrastructure only.
"""
from __future__ import annotations

import copy

import torch
import torch.nn as nn
import torch.nn.functional as F


class ControlNetConditioningEmbedding(nn.Module):
    def __init__(self, conditioning_embedding_channels, conditioning_channels=3, block_out_channels=(16, 32, 96, 256)):
        super().__init__()
        boc = tuple(block_out_channels)
        self.conv_in = nn.Conv2d(conditioning_channels, boc[0], 3, padding=1)
        self.blocks = nn.ModuleList()
        for i in range(len(boc) - 1):
            self.blocks.append(nn.Conv2d(boc[i], boc[i], 3, padding=1))
            self.blocks.append(nn.Conv2d(boc[i], boc[i + 1], 3, padding=1, stride=2))
        self.conv_out = nn.Conv2d(boc[-1], conditioning_embedding_channels, 3, padding=1)
        nn.init.zeros_(self.conv_out.weight)
        nn.init.zeros_(self.conv_out.bias)

    def forward(self, conditioning):
        e = F.silu(self.conv_in(conditioning))
        for blk in self.blocks:
            e = F.silu(blk(e))
        return self.conv_out(e)


def _zero_conv1x1(c):
    conv = nn.Conv2d(c, c, 1)
    nn.init.zeros_(conv.weight)
    nn.init.zeros_(conv.bias)
    return conv


class ControlNetModel(nn.Module):
    """Attribute tree = diffusers': time_proj, time_embedding, conv_in, controlnet_cond_embedding, down_blocks,
    mid_block, controlnet_down_blocks, controlnet_mid_block."""

    def __init__(self, unet, conditioning_embedding_out_channels=(16, 32, 96, 256)):
        super().__init__()
        self.config = unet.config
        self.conv_in = copy.deepcopy(unet.conv_in)
        self.time_proj = copy.deepcopy(unet.time_proj)
        self.time_embedding = copy.deepcopy(unet.time_embedding)
        self.controlnet_cond_embedding = ControlNetConditioningEmbedding(
            unet.config.block_out_channels[0], block_out_channels=conditioning_embedding_out_channels)
        self.down_blocks = copy.deepcopy(unet.down_blocks)
        self.mid_block = copy.deepcopy(unet.mid_block)
        chans = [unet.config.block_out_channels[0]]
        for blk in unet.down_blocks:
            chans += [r.out_channels for r in blk.resnets]
            if blk.downsamplers is not None:
                chans.append(blk.downsamplers[-1].out_channels)
        self.controlnet_down_blocks = nn.ModuleList([_zero_conv1x1(c) for c in chans])
        self.controlnet_mid_block = _zero_conv1x1(unet.mid_block.resnets[-1].out_channels)

    @classmethod
    def from_unet(cls, unet, **kw):
        """Weights of conv_in / time embedding / down / mid are copied from the UNet (load_weights_from_unet=True)."""
        return cls(unet, **kw)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    def forward(self, sample, timestep, encoder_hidden_states, controlnet_cond, conditioning_scale=1.0,
                return_dict=False):
        t = timestep
        if t.dim() == 0:
            t = t[None]
        t = t.expand(sample.shape[0]) if t.shape[0] != sample.shape[0] else t
        emb = self.time_embedding(self.time_proj(t).to(sample.dtype))
        sample = self.conv_in(sample) + self.controlnet_cond_embedding(controlnet_cond)
        res = (sample,)
        for blk in self.down_blocks:
            attn = getattr(blk, "has_cross_attention", False)
            for j, r in enumerate(blk.resnets):
                sample = r(sample, emb)
                if attn:
                    sample = blk.attentions[j](sample, encoder_hidden_states=encoder_hidden_states).sample
                res += (sample,)
            if blk.downsamplers is not None:
                for d in blk.downsamplers:
                    sample = d(sample)
                res += (sample,)
        sample = self.mid_block.resnets[0](sample, emb)
        for i, a in enumerate(self.mid_block.attentions):
            sample = a(sample, encoder_hidden_states=encoder_hidden_states).sample
            sample = self.mid_block.resnets[i + 1](sample, emb)
        down = tuple(conv(s) * conditioning_scale for s, conv in zip(res, self.controlnet_down_blocks))
        mid = self.controlnet_mid_block(sample) * conditioning_scale
        assert not return_dict
        return down, mid


def randomize_zero_convs(cn: ControlNetModel, seed: int = 4321, std: float = 0.02) -> None:
    """The zero-initialised convs make a fresh ControlNet contribute nothing (bugs invisible): redraw every all-zero
    parameter from N(0, std^2), deterministically ('zero-convs re-drawn N(0,0.02^2)')."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for _, p in sorted(cn.named_parameters()):
            if p.abs().sum() == 0:
                p.copy_(torch.randn(p.shape, generator=g) * std)


def build_controlnet(unet, seed: int = 7) -> ControlNetModel:
    """from_unet copy; the cond-embedding convs get PyTorch default init under `seed`, zero convs are redrawn."""
    st = torch.random.get_rng_state()
    torch.manual_seed(seed)
    cn = ControlNetModel.from_unet(unet).eval()
    torch.random.set_rng_state(st)
    randomize_zero_convs(cn, seed + 1)
    return cn
