"""Oracle: the forward half of the training step — restates models/pano/PanFusion.py:78-97 (timestep, shared noise field,
add_noise on both latents, joint denoiser forward, the two MSE terms). diffusers `DDIMScheduler.add_noise` is third-party and
absent here: restated from its published rule, noisy = sqrt(abar_t) x0 + sqrt(1 - abar_t) eps [3P, parity unpinned]; the
denoiser is the oracle model pinned by the reference goldens. Test infrastructure only."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import sampler


def add_noise(x0, noise, t, alphas_cumprod):
    """diffusers SchedulerMixin.add_noise: per-sample coefficients broadcast over the trailing dims."""
    a = alphas_cumprod.to(x0.dtype)[t]
    sa, s1 = a ** 0.5, (1 - a) ** 0.5
    shape = (-1,) + (1,) * (x0.dim() - 1)
    return sa.reshape(shape) * x0 + s1.reshape(shape) * noise


def training_loss(model, latents, pano_latent, t, pers_prompt_embd, pano_prompt_embd, cameras, pano_noise):
    """PanFusion.py:81-94 with the random draws (t, pano_noise) supplied."""
    b, m = latents.shape[:2]
    ddim = sampler.DDIM()
    noise = sampler.init_noise(pano_noise, latents.shape[-2], latents.shape[-1], cameras)
    noise_z = add_noise(latents, noise, t, ddim.alphas_cumprod)
    pano_noise_z = add_noise(pano_latent, pano_noise, t, ddim.alphas_cumprod)
    tm = t[:, None].repeat(1, m)
    with torch.no_grad():
        denoise, pano_denoise = model(noise_z, pano_noise_z, tm, pers_prompt_embd, pano_prompt_embd, cameras)
    loss_pers = F.mse_loss(denoise, noise)
    loss_pano = F.mse_loss(pano_denoise, pano_noise)
    return dict(loss=loss_pers + loss_pano, loss_pers=loss_pers, loss_pano=loss_pano, noise=noise, noise_z=noise_z,
                pano_noise_z=pano_noise_z)
