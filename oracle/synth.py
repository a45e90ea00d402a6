"""Seeded synthetic inputs / weights shared by the golden generator, the tests and bench.py.
Test infrastructure only."""
from __future__ import annotations

import torch

from . import controlnet as ocn, sampler, unet as ounet


def randomize_zero_init(model: torch.nn.Module, seed: int = 1234, std: float = 0.02) -> None:
    """EPPA output projections are zero-initialised (transformer.py:29-30,54-55), which makes a fresh WarpAttn the
    identity and hides bugs: redraw every all-zero parameter of the cp_blocks from N(0, std^2), deterministically."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(model.named_parameters()):
            if "cp_blocks" in name and p.abs().sum() == 0:
                p.copy_(torch.randn(p.shape, generator=g) * std)


def step_inputs(m: int, pano_hw, pers_hw, ctx_dim: int, seed: int = 0, batch: int = 1, t: int = 981):
    """Inputs of ONE MultiViewBaseModel.forward call (no CFG duplication): latents via init_noise's shared field."""
    g = torch.Generator().manual_seed(seed)
    cams = sampler.horizon_cameras(m, batch=batch)
    pano = torch.randn(batch, 1, 4, *pano_hw, generator=g)
    lat = sampler.init_noise(pano, pers_hw[0], pers_hw[1], cams)
    ts = torch.full((batch, m), t, dtype=torch.long)
    prompt = torch.randn(batch, 1, 77, ctx_dim, generator=g).repeat(1, m, 1, 1)  # copy_pano_prompt (PanFusion.py:17)
    pano_prompt = prompt[:, :1].clone()
    return dict(latents=lat, pano_latent=pano, timestep=ts, prompt_embd=prompt, pano_prompt_embd=pano_prompt,
                cameras=cams)


def step_inputs_cfg(m: int, pano_hw, pers_hw, ctx_dim: int, seed: int = 0, t: int = 981):
    """Inputs of the CFG-batched forward exactly as PanFusion.forward_cls_free hands them to the denoiser
    (PanFusion.py:100-108, PanoGenerator.py:240-251): every input duplicated along the batch, prompts = [null; text]
    (PanFusion.py:135-138), cameras duplicated. b = 2, m views (BASELINE configs[1] for m = 8)."""
    g = torch.Generator().manual_seed(seed)
    cams = sampler.horizon_cameras(m, batch=1)
    pano = torch.randn(1, 1, 4, *pano_hw, generator=g)
    lat = sampler.init_noise(pano, pers_hw[0], pers_hw[1], cams)
    text = torch.randn(1, 1, 77, ctx_dim, generator=g)
    null = torch.randn(1, 1, 77, ctx_dim, generator=g)
    dup = lambda x: torch.cat([x] * 2)
    return dict(latents=dup(lat), pano_latent=dup(pano), timestep=torch.full((2, m), t, dtype=torch.long),
                prompt_embd=torch.cat([null.repeat(1, m, 1, 1), text.repeat(1, m, 1, 1)]),
                pano_prompt_embd=torch.cat([null, text]), cameras={k: dup(v) for k, v in cams.items()})


def build_model(model_cls, config=None, seed: int = 0):
    """Two independently seeded UNets + the EPPA blocks with their zero-init tensors redrawn."""
    config = config or ounet.SD2_CONFIG
    st = torch.random.get_rng_state()
    torch.manual_seed(seed + 100)
    unet = ounet.build_unet(config, seed=seed + 1)
    pano_unet = ounet.build_unet(config, seed=seed + 2)
    model = model_cls(unet, pano_unet).eval()
    torch.random.set_rng_state(st)
    randomize_zero_init(model, seed + 3)
    return model


def build_model_cn(model_cls, config=None, seed: int = 0, pers: bool = False):
    """build_model + a panorama ControlNet (and optionally a perspective one; off by default in the reference,
    PanoGenerator.py:77). The ControlNet encoders are `from_unet` copies of DIFFERENTLY seeded UNets, so using the
    UNet's weights in place of the ControlNet's shows up in the output."""
    config = config or ounet.SD2_CONFIG
    base = build_model(model_cls, config, seed)
    pano_cn = ocn.build_controlnet(ounet.build_unet(config, seed=seed + 11), seed=seed + 12)
    pers_cn = ocn.build_controlnet(ounet.build_unet(config, seed=seed + 13), seed=seed + 14) if pers else None
    model = model_cls(base.unet, base.pano_unet, pers_cn=pers_cn, pano_cn=pano_cn).eval()
    model.load_state_dict(base.state_dict(), strict=False)  # the EPPA blocks (incl. redrawn zero-init tensors)
    return model


def layout_conds(batch: int, m: int, pano_hw, pers_hw, seed: int = 5, pers: bool = False):
    """Layout-condition images at 8x the latent resolution (PanFusion.py:152-153 rolls the pano one 256 px/step)."""
    g = torch.Generator().manual_seed(seed)
    out = dict(pano_layout_cond=torch.rand(batch, 1, 3, pano_hw[0] * 8, pano_hw[1] * 8, generator=g))
    if pers:
        out["pers_layout_cond"] = torch.rand(batch, m, 3, pers_hw[0] * 8, pers_hw[1] * 8, generator=g)
    return out
