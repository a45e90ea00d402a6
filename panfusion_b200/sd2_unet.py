"""Parameter tree of the Stable-Diffusion-2 `UNet2DConditionModel` (diffusers 0.24.0 attribute / state-dict names),
WITHOUT any forward pass: `MultiViewBaseModel` only reads parameters from it (engine.UNetPack), exactly like the
reference only ever touches the UNet through attributes (models/pano/MVGenModel.py:52-295). Lets the framework be
instantiated, loaded from a converted checkpoint and benchmarked where `diffusers` is not installed.
The real diffusers module (or any module with the same attribute tree) can be passed instead.
"""
from __future__ import annotations

import torch
import torch.nn as nn

SD2_CONFIG = dict(
    in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
    attention_heads=(5, 10, 20, 20), cross_attention_dim=1024, norm_num_groups=32,
    down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
)


class _Holder(nn.Module):
    """A named bag of sub-modules / attributes; calling it is an error (there is no PyTorch compute path)."""

    def forward(self, *a, **k):
        raise RuntimeError("panfusion_b200.sd2_unet holds parameters only; run it through MultiViewBaseModel")


def _resnet(cin, cout, temb, groups):
    r = _Holder()
    r.in_channels, r.out_channels = cin, cout
    r.norm1 = nn.GroupNorm(groups, cin, eps=1e-5)
    r.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
    r.time_emb_proj = nn.Linear(temb, cout)
    r.norm2 = nn.GroupNorm(groups, cout, eps=1e-5)
    r.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
    r.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None
    return r


def _attention(dim, heads, ctx=None):
    a = _Holder()
    a.heads = heads
    a.to_q = nn.Linear(dim, dim, bias=False)
    a.to_k = nn.Linear(ctx or dim, dim, bias=False)
    a.to_v = nn.Linear(ctx or dim, dim, bias=False)
    a.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])
    return a


def _transformer(dim, heads, ctx, groups):
    t = _Holder()
    t.norm = nn.GroupNorm(groups, dim, eps=1e-6)
    t.proj_in = nn.Linear(dim, dim)
    blk = _Holder()
    blk.norm1, blk.attn1 = nn.LayerNorm(dim), _attention(dim, heads)
    blk.norm2, blk.attn2 = nn.LayerNorm(dim), _attention(dim, heads, ctx)
    blk.norm3 = nn.LayerNorm(dim)
    ff = _Holder()
    geglu = _Holder()
    geglu.proj = nn.Linear(dim, dim * 8)
    ff.net = nn.ModuleList([geglu, nn.Dropout(0.0), nn.Linear(dim * 4, dim)])
    blk.ff = ff
    t.transformer_blocks = nn.ModuleList([blk])
    t.proj_out = nn.Linear(dim, dim)
    return t


def _sampler(channels, stride):
    s = _Holder()
    s.channels = s.out_channels = channels
    s.conv = nn.Conv2d(channels, channels, 3, stride=stride, padding=1)
    return s


class SD2UNetParams(_Holder):
    def __init__(self, in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                 attention_heads=(5, 10, 20, 20), cross_attention_dim=1024, norm_num_groups=32,
                 down_block_types=SD2_CONFIG["down_block_types"], up_block_types=SD2_CONFIG["up_block_types"]):
        super().__init__()
        boc, g, ctx = tuple(block_out_channels), norm_num_groups, cross_attention_dim
        temb = boc[0] * 4
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_proj = _Holder()
        self.time_proj.num_channels = boc[0]
        te = _Holder()
        te.linear_1, te.linear_2 = nn.Linear(boc[0], temb), nn.Linear(temb, temb)
        self.time_embedding = te
        self.down_blocks = nn.ModuleList()
        out = boc[0]
        for i, typ in enumerate(down_block_types):
            cin, out = out, boc[i]
            b = _Holder()
            b.has_cross_attention = typ.startswith("CrossAttn")
            b.resnets = nn.ModuleList([_resnet(cin if j == 0 else out, out, temb, g) for j in range(layers_per_block)])
            if b.has_cross_attention:
                b.attentions = nn.ModuleList([_transformer(out, attention_heads[i], ctx, g) for _ in range(layers_per_block)])
            b.downsamplers = nn.ModuleList([_sampler(out, 2)]) if i != len(boc) - 1 else None
            self.down_blocks.append(b)
        mid = _Holder()
        mid.has_cross_attention = True
        mid.resnets = nn.ModuleList([_resnet(boc[-1], boc[-1], temb, g) for _ in range(2)])
        mid.attentions = nn.ModuleList([_transformer(boc[-1], attention_heads[-1], ctx, g)])
        self.mid_block = mid
        self.up_blocks = nn.ModuleList()
        rev, rev_heads = boc[::-1], tuple(attention_heads)[::-1]
        out = rev[0]
        n_layers = layers_per_block + 1
        for i, typ in enumerate(up_block_types):
            prev, out = out, rev[i]
            cin = rev[min(i + 1, len(boc) - 1)]
            b = _Holder()
            b.has_cross_attention = typ.startswith("CrossAttn")
            b.resnets = nn.ModuleList([
                _resnet((prev if j == 0 else out) + (cin if j == n_layers - 1 else out), out, temb, g)
                for j in range(n_layers)])
            if b.has_cross_attention:
                b.attentions = nn.ModuleList([_transformer(out, rev_heads[i], ctx, g) for _ in range(n_layers)])
            b.upsamplers = nn.ModuleList([_sampler(out, 1)]) if i != len(boc) - 1 else None
            self.up_blocks.append(b)
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-5)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype


def build_synthetic(config: dict | None = None, seed: int = 0, device="cpu") -> SD2UNetParams:
    """Random-init weights of the SD-2 architecture (PyTorch default init under `seed`), created on `device`."""
    config = config or SD2_CONFIG
    dev = torch.device(device)
    gens = torch.random.get_rng_state()
    torch.manual_seed(seed)
    if dev.type == "cuda":
        torch.cuda.manual_seed(seed)
    with torch.device(dev):
        net = SD2UNetParams(**config)
    torch.random.set_rng_state(gens)
    return net.eval()


class SD2ControlNetParams(_Holder):
    """Parameter tree of diffusers `ControlNetModel.from_unet(unet)` [3P] (models/pano/PanoGenerator.py:153-157):
    `conv_in`, `time_proj`, `time_embedding`, `down_blocks`, `mid_block` (same attribute tree as the UNet encoder),
    `controlnet_cond_embedding.{conv_in, blocks[6], conv_out}` (3x3 convs 3->16->16->32->32->96->96->256->C0, strides
    1,1,2,1,2,1,2,1), `controlnet_down_blocks[12]` and `controlnet_mid_block` (1x1 convs)."""

    def __init__(self, conditioning_embedding_out_channels=(16, 32, 96, 256), **config):
        super().__init__()
        enc = SD2UNetParams(**config)
        self.conv_in, self.time_proj, self.time_embedding = enc.conv_in, enc.time_proj, enc.time_embedding
        self.down_blocks, self.mid_block = enc.down_blocks, enc.mid_block
        boc = tuple(conditioning_embedding_out_channels)
        ce = _Holder()
        ce.conv_in = nn.Conv2d(3, boc[0], 3, padding=1)
        ce.blocks = nn.ModuleList()
        for i in range(len(boc) - 1):
            ce.blocks.append(nn.Conv2d(boc[i], boc[i], 3, padding=1))
            ce.blocks.append(nn.Conv2d(boc[i], boc[i + 1], 3, padding=1, stride=2))
        c0 = enc.conv_in.out_channels
        ce.conv_out = nn.Conv2d(boc[-1], c0, 3, padding=1)
        self.controlnet_cond_embedding = ce
        chans = [c0]
        for blk in self.down_blocks:
            chans += [r.out_channels for r in blk.resnets]
            if blk.downsamplers is not None:
                chans.append(blk.downsamplers[-1].out_channels)
        self.controlnet_down_blocks = nn.ModuleList([nn.Conv2d(c, c, 1) for c in chans])
        cm = self.mid_block.resnets[-1].out_channels
        self.controlnet_mid_block = nn.Conv2d(cm, cm, 1)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype


def build_synthetic_controlnet(config: dict | None = None, seed: int = 0, device="cpu") -> SD2ControlNetParams:
    """Random-init ControlNet of the SD-2 architecture (PyTorch default init under `seed`; the zero-initialised convs of
    a fresh diffusers ControlNet are drawn like every other conv so that they contribute to the output)."""
    config = config or SD2_CONFIG
    dev = torch.device(device)
    gens = torch.random.get_rng_state()
    torch.manual_seed(seed)
    if dev.type == "cuda":
        torch.cuda.manual_seed(seed)
    with torch.device(dev):
        net = SD2ControlNetParams(**config)
    torch.random.set_rng_state(gens)
    return net.eval()


# ---- VAE decoder (image-space tail of the loop) ------------------------------------------

SD2_VAE_CONFIG = dict(latent_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                      norm_num_groups=32, scaling_factor=0.18215)


def _vae_resnet(cin, cout, groups):
    r = _Holder()
    r.in_channels, r.out_channels = cin, cout
    r.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
    r.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
    r.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
    r.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
    r.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None
    return r


class SD2VAEDecoderParams(_Holder):
    """Parameter tree of the decoder half of diffusers `AutoencoderKL` [3P] (`post_quant_conv`, `decoder.*`, state-dict
    names as in `stabilityai/stable-diffusion-2-base/vae`), consumed by panfusion_b200.vae.VAEDecoder."""

    def __init__(self, latent_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 norm_num_groups=32, scaling_factor=0.18215):
        super().__init__()
        from types import SimpleNamespace
        boc, g = tuple(block_out_channels), norm_num_groups
        self.config = SimpleNamespace(latent_channels=latent_channels, out_channels=out_channels, block_out_channels=boc,
                                      layers_per_block=layers_per_block, norm_num_groups=g,
                                      scaling_factor=scaling_factor)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)
        d = _Holder()
        d.conv_in = nn.Conv2d(latent_channels, boc[-1], 3, padding=1)
        mid = _Holder()
        mid.resnets = nn.ModuleList([_vae_resnet(boc[-1], boc[-1], g) for _ in range(2)])
        att = _Holder()
        att.heads = 1
        att.group_norm = nn.GroupNorm(g, boc[-1], eps=1e-6)
        att.to_q, att.to_k, att.to_v = (nn.Linear(boc[-1], boc[-1]) for _ in range(3))
        att.to_out = nn.ModuleList([nn.Linear(boc[-1], boc[-1]), nn.Dropout(0.0)])
        mid.attentions = nn.ModuleList([att])
        d.mid_block = mid
        d.up_blocks = nn.ModuleList()
        rev = boc[::-1]
        out = rev[0]
        for i, c in enumerate(rev):
            prev, out = out, c
            b = _Holder()
            b.resnets = nn.ModuleList([_vae_resnet(prev if j == 0 else out, out, g) for j in range(layers_per_block + 1)])
            b.upsamplers = nn.ModuleList([_sampler(out, 1)]) if i != len(rev) - 1 else None
            d.up_blocks.append(b)
        d.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-6)
        d.conv_act = nn.SiLU()
        d.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)
        self.decoder = d

    @property
    def dtype(self):
        return self.post_quant_conv.weight.dtype


def build_synthetic_vae(config: dict | None = None, seed: int = 0, device="cpu") -> SD2VAEDecoderParams:
    config = config or SD2_VAE_CONFIG
    dev = torch.device(device)
    gens = torch.random.get_rng_state()
    torch.manual_seed(seed)
    if dev.type == "cuda":
        torch.cuda.manual_seed(seed)
    with torch.device(dev):
        net = SD2VAEDecoderParams(**config)
    torch.random.set_rng_state(gens)
    return net.eval()
