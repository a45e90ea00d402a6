"""Oracle: the dual-branch denoiser `MultiViewBaseModel` in plain PyTorch fp32.

Restates models/pano/MVGenModel.py:8-297: perspective UNet and panorama UNet walked block by block in lock-step,
every panorama convolution wrapped in circular pad / unpad (utils/pano.py:74-105), seven EPPA fusions
(encoder after each downsampler, mid, decoder before each upsampler). ControlNet residual hooks
(MVGenModel.py:62-83,154-170,200-203) are kept. Test infrastructure only.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .eppa import WarpAttn, pad_pano, unpad_pano


class MultiViewBaseModel(nn.Module):
    def __init__(self, unet, pano_unet, pers_cn=None, pano_cn=None, pano_pad=True):
        super().__init__()
        self.unet, self.pano_unet = unet, pano_unet
        self.pers_cn, self.pano_cn, self.pano_pad = pers_cn, pano_cn, pano_pad
        if unet is not None:  # MVGenModel.py:17-36
            self.cp_blocks_encoder = nn.ModuleList(
                [WarpAttn(blk.downsamplers[-1].out_channels) for blk in unet.down_blocks if blk.downsamplers is not None])
            self.cp_blocks_mid = WarpAttn(unet.mid_block.resnets[-1].out_channels)
            self.cp_blocks_decoder = nn.ModuleList(
                [WarpAttn(blk.upsamplers[0].channels) for blk in unet.up_blocks if blk.upsamplers is not None])
            self.trainable_parameters = [(list(self.cp_blocks_mid.parameters())
                                          + list(self.cp_blocks_decoder.parameters())
                                          + list(self.cp_blocks_encoder.parameters()), 1.0)]

    # panorama-side helper: run `fn` on the circularly padded tensor, crop `out_pad` columns per side afterwards
    def _pano(self, fn, x, in_pad, out_pad=None):
        out_pad = in_pad if out_pad is None else out_pad
        if not self.pano_pad:
            return fn(x)
        return unpad_pano(fn(pad_pano(x, in_pad)), out_pad)

    def forward(self, latents, pano_latent, timestep, prompt_embd, pano_prompt_embd, cameras,
                pers_layout_cond=None, pano_layout_cond=None):
        pers = self.unet is not None
        if latents is not None:
            b, m = latents.shape[:2]
            h = latents.flatten(0, 1)
        if cameras is not None:
            cameras = {k: v.flatten(0, 1) for k, v in cameras.items()}
        if prompt_embd is not None:
            prompt_embd = prompt_embd.flatten(0, 1)
        p = pano_latent.flatten(0, 1)
        pano_prompt_embd = pano_prompt_embd.flatten(0, 1)

        # timesteps (MVGenModel.py:52-60): pers uses every (b, m) entry, pano the first view's
        if pers:
            pano_t = timestep[:, 0]
            t = timestep.reshape(-1)
            emb = self.unet.time_embedding(self.unet.time_proj(t).to(self.unet.dtype))
        else:
            pano_t = timestep
        pano_emb = self.pano_unet.time_embedding(self.pano_unet.time_proj(pano_t).to(self.pano_unet.dtype))

        # ControlNets (MVGenModel.py:62-83); pano ControlNet sees the UNPADDED latent
        pers_cn_res = pano_cn_res = None
        if self.pers_cn is not None and pers_layout_cond is not None:
            pers_cn_res = self.pers_cn(h, t, encoder_hidden_states=prompt_embd,
                                       controlnet_cond=pers_layout_cond.flatten(0, 1), return_dict=False)
        if self.pano_cn is not None and pano_layout_cond is not None:
            pano_cn_res = self.pano_cn(p, pano_t, encoder_hidden_states=pano_prompt_embd,
                                       controlnet_cond=pano_layout_cond.flatten(0, 1), return_dict=False)

        # conv_in (MVGenModel.py:85-91)
        if pers:
            h = self.unet.conv_in(h)
        p = self._pano(self.pano_unet.conv_in, p, 1)

        skips, pano_skips = ([h] if pers else []), [p]
        # encoder (MVGenModel.py:98-152)
        enc = 0
        for i, pblk in enumerate(self.pano_unet.down_blocks):
            has_attn = getattr(pblk, "has_cross_attention", False)
            for j in range(len(pblk.resnets)):
                if pers:
                    blk = self.unet.down_blocks[i]
                    h = blk.resnets[j](h, emb)
                    if has_attn:
                        h = blk.attentions[j](h, encoder_hidden_states=prompt_embd).sample
                    skips.append(h)
                p = self._pano(lambda x: pblk.resnets[j](x, pano_emb), p, 2)
                if has_attn:
                    p = pblk.attentions[j](p, encoder_hidden_states=pano_prompt_embd).sample
                pano_skips.append(p)
            if pblk.downsamplers is not None:
                for j in range(len(pblk.downsamplers)):
                    if pers:
                        h = self.unet.down_blocks[i].downsamplers[j](h)
                    p = self._pano(pblk.downsamplers[j], p, 2, 1)
                if pers:
                    skips.append(h)
                pano_skips.append(p)  # skip saved BEFORE the fusion
                if pers:
                    h, p = self.cp_blocks_encoder[i](h, p, cameras)

        if pers_cn_res is not None:
            skips = [s + r for s, r in zip(skips, pers_cn_res[0])]
        if pano_cn_res is not None:
            pano_skips = [s + r for s, r in zip(pano_skips, pano_cn_res[0])]

        # mid (MVGenModel.py:172-207)
        pmid = self.pano_unet.mid_block
        if pers:
            h = self.unet.mid_block.resnets[0](h, emb)
        p = self._pano(lambda x: pmid.resnets[0](x, pano_emb), p, 2)
        for i in range(len(pmid.attentions)):
            if pers:
                h = self.unet.mid_block.attentions[i](h, encoder_hidden_states=prompt_embd).sample
                h = self.unet.mid_block.resnets[i + 1](h, emb)
            p = pmid.attentions[i](p, encoder_hidden_states=pano_prompt_embd).sample
            p = self._pano(lambda x: pmid.resnets[i + 1](x, pano_emb), p, 2)
        if pers_cn_res is not None:
            h = h + pers_cn_res[1]
        if pano_cn_res is not None:
            p = p + pano_cn_res[1]
        if pers:
            h, p = self.cp_blocks_mid(h, p, cameras)

        # decoder (MVGenModel.py:210-277)
        for i, pblk in enumerate(self.pano_unet.up_blocks):
            has_attn = getattr(pblk, "has_cross_attention", False)
            for j in range(len(pblk.resnets)):
                if pers:
                    blk = self.unet.up_blocks[i]
                    h = blk.resnets[j](torch.cat([h, skips.pop()], dim=1), emb)
                    if has_attn:
                        h = blk.attentions[j](h, encoder_hidden_states=prompt_embd).sample
                p = torch.cat([p, pano_skips.pop()], dim=1)
                p = self._pano(lambda x: pblk.resnets[j](x, pano_emb), p, 2)
                if has_attn:
                    p = pblk.attentions[j](p, encoder_hidden_states=pano_prompt_embd).sample
            if pblk.upsamplers is not None:
                if pers:
                    h, p = self.cp_blocks_decoder[i](h, p, cameras)  # fusion BEFORE the upsampler
                for j in range(len(pblk.upsamplers)):
                    if pers:
                        h = self.unet.up_blocks[i].upsamplers[j](h)
                    p = self._pano(pblk.upsamplers[j], p, 1, 2)

        # output heads (MVGenModel.py:279-297)
        sample = None
        if pers:
            sample = self.unet.conv_out(self.unet.conv_act(self.unet.conv_norm_out(h)))
            sample = sample.reshape(b, m, *sample.shape[1:])
        p = self.pano_unet.conv_act(self.pano_unet.conv_norm_out(p))
        p = self._pano(self.pano_unet.conv_out, p, 1)
        return sample, p[:, None]
