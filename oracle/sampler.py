"""Oracle: the sampling loop around the denoiser — rotation, classifier-free guidance, DDIM update.

Restates models/pano/PanFusion.py:30-43 (init_noise), :100-123 (forward_cls_free, rotate_latent), :146-162 (the hot
loop) and models/pano/PanoGenerator.py:240-269 (CFG pair / combine, latent roll). diffusers DDIMScheduler
(SD-2 config: 1000 train steps, scaled_linear 0.00085..0.012, steps_offset 1, epsilon prediction, eta 0,
set_alpha_to_one False, 'leading' spacing) is restated from its published update rule [3P, parity unpinned].
Test infrastructure only.
"""
from __future__ import annotations

import numpy as np
import torch

from . import geometry as G


class DDIM:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = self.alphas_cumprod[0]
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.timesteps = None
        self.num_inference_steps = None

    def set_timesteps(self, n):
        self.num_inference_steps = n
        ratio = self.num_train_timesteps // n
        ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset
        self.timesteps = torch.from_numpy(ts)

    def step(self, eps, t, x):
        t = int(t)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        return a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * eps


def horizon_cameras(m, fov=90.0, batch=1):
    """utils/pano.py:28-31 (theta = 0..360 step 360/m, phi = 0) in degrees, as dataset/PanoDataset.py:117-126
    stores them: tensors [batch, m]."""
    theta = np.rad2deg(np.linspace(0, 2 * np.pi, m, endpoint=False))
    mk = lambda v: torch.tensor(v, dtype=torch.float32)[None].repeat(batch, 1)
    return {"FoV": mk(np.full(m, fov)), "theta": mk(theta), "phi": mk(np.zeros(m))}


def init_noise(pano_noise, pers_h, pers_w, cameras):
    """PanFusion.py:30-43 with the panorama noise supplied: every view's noise is the nearest-neighbour e2p
    resampling of the shared panorama noise field. pano_noise [bs,1,4,H,W] -> noise [bs,m,4,h,w]."""
    bs = pano_noise.shape[0]
    cams = {k: v.flatten(0, 1) for k, v in cameras.items()}
    m = len(cams["FoV"]) // bs
    rep = pano_noise.expand(-1, m, -1, -1, -1).flatten(0, 1)
    noise = G.e2p(rep, cams["FoV"], cams["theta"], cams["phi"], (pers_h, pers_w), mode="nearest")
    return noise.reshape(bs, m, *noise.shape[1:])


def rotate_latent(pano_latent, cameras, degree=90.0):
    """PanFusion.py:114-123 + PanoGenerator.py:264-269."""
    if degree % 360 == 0:
        return pano_latent, cameras
    pano_latent = torch.roll(pano_latent, int(degree / 360 * pano_latent.shape[-1]), dims=-1)
    cameras = dict(cameras)
    cameras["theta"] = (cameras["theta"] + degree) % 360
    return pano_latent, cameras


def forward_cls_free(model, latents, pano_latent, timestep, prompt_embd, pano_prompt_embd, cameras,
                     guidance_scale=9.0, pers_layout_cond=None, pano_layout_cond=None):
    """PanFusion.py:100-112: duplicate inputs (uncond first, text second; layout conditions too,
    PanoGenerator.py:240-251), one forward, combine."""
    dup = lambda t: torch.cat([t] * 2) if t is not None else None
    cams2 = {k: dup(v) for k, v in cameras.items()}
    eps, pano_eps = model(dup(latents), dup(pano_latent), dup(timestep), prompt_embd, pano_prompt_embd, cams2,
                          dup(pers_layout_cond), dup(pano_layout_cond))
    def combine(e):
        u, c = e.chunk(2)
        return u + guidance_scale * (c - u)
    return combine(eps), combine(pano_eps)


@torch.no_grad()
def denoise_steps(model, latents, pano_latent, prompt_embd, pano_prompt_embd, cameras, num_steps,
                  diff_timestep=50, guidance_scale=9.0, rot_diff=90.0, start_step=0, pers_layout_cond=None,
                  pano_layout_cond=None):
    """`num_steps` iterations of the hot loop PanFusion.py:146-162. prompt_embd / pano_prompt_embd are the CFG
    concatenations [null; text] of shape [2, m, 77, C] / [2, 1, 77, C]. Returns (latents, pano_latent, cameras)."""
    sched = DDIM()
    sched.set_timesteps(diff_timestep)
    m = latents.shape[1]
    for t in sched.timesteps[start_step:start_step + num_steps]:
        timestep = torch.cat([t[None, None]] * m, dim=1)
        pano_latent, cameras = rotate_latent(pano_latent, cameras, rot_diff)
        if pano_layout_cond is not None and rot_diff % 360:  # PanFusion.py:152-153: rolled every step, cumulatively
            pano_layout_cond = torch.roll(pano_layout_cond, int(rot_diff / 360 * pano_layout_cond.shape[-1]), dims=-1)
        eps, pano_eps = forward_cls_free(model, latents, pano_latent, timestep, prompt_embd, pano_prompt_embd,
                                         cameras, guidance_scale, pers_layout_cond, pano_layout_cond)
        latents = sched.step(eps, t, latents)
        pano_latent = sched.step(pano_eps, t, pano_latent)
    return latents, pano_latent, cameras
