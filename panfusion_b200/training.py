"""Forward half of the reference's training step: `PanFusion.training_step` (models/pano/PanFusion.py:64-98) from the sampled
timestep to the loss — `init_noise` (PanFusion.py:30-43), `scheduler.add_noise` on both latents (:84-85, diffusers
`DDIMScheduler.add_noise` [3P]), the joint denoiser forward (:88-90) and the two epsilon-prediction MSE terms (:92-94) — on
the CUDA kernels of the denoise path plus `pf_add_noise` / `pf_mse_loss`.

SURVEY.md 8f rank 4 names the whole training step; what is NOT built is everything after the loss: the backward of both
UNets, of EPPA and of the LoRA adapters, the optimizer and the gradient all-reduce. The clean latents come from
`TrainingStep.encode` (`encode_image` on the views and on the circularly padded panorama, PanFusion.py:66-71, on
`panfusion_b200.vae.VAEEncoder`) or from the caller. `TrainingStep.loss` is therefore a validation /
monitoring quantity (the number the reference logs as train/loss), not a trainable graph: the returned tensors carry no
autograd history and `backward()` raises.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import ops
from .sampler import DDIMSchedule, init_noise


class TrainingStep:
    def __init__(self, mv_base_model, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012):
        self.mv_base_model = mv_base_model
        self.scheduler = DDIMSchedule(num_train_timesteps, beta_start, beta_end)  # PanoGenerator.py:129-130 (SD-2 config)
        self.num_train_timesteps = num_train_timesteps
        self._abar = {}

    def _alphas_cumprod(self, device) -> Tensor:
        a = self._abar.get(device)
        if a is None:
            a = self._abar[device] = self.scheduler.alphas_cumprod.to(device=device, dtype=torch.float32).contiguous()
        return a

    @staticmethod
    def encode(images: Tensor, pano: Tensor, vae, latent_pad: int = 8, generator=None):
        """PanFusion.py:66-71: images [b, m, 3, H, W], pano [b, 1, 3, He, We] in [-1, 1] -> (latents, pano_latent); the
        panorama is padded circularly in IMAGE space by 8 * latent_pad pixels, encoded, and cropped in latent space.
        vae: panfusion_b200.vae.VAEEncoder."""
        from .vae import encode_image, encode_pano
        return encode_image(images, vae, generator=generator), encode_pano(pano, vae, latent_pad, generator=generator)

    def add_noise(self, x0: Tensor, noise: Tensor, t: Tensor) -> Tensor:
        """scheduler.add_noise(x0, noise, t) with one timestep per leading-dim sample (PanFusion.py:84-85)."""
        return ops.add_noise(x0.float().contiguous(), noise.float().contiguous(), t.to(torch.int64).contiguous(),
                             self._alphas_cumprod(x0.device))

    @torch.no_grad()
    def loss(self, latents: Tensor, pano_latent: Tensor, pers_prompt_embd: Tensor, pano_prompt_embd: Tensor, cameras: dict,
             t: Optional[Tensor] = None, noise: Optional[Tensor] = None, pano_noise: Optional[Tensor] = None,
             generator=None, images_layout_cond=None, pano_layout_cond=None) -> dict:
        """latents [b, m, 4, h, w], pano_latent [b, 1, 4, H, W] (the VAE-encoded batch, PanFusion.py:66-71), prompts as
        `embed_prompt` returns them, cameras {FoV, theta, phi: [b, m]}. t / noise / pano_noise default to the reference's
        random draws (:78-83); pass them to reproduce a given step. -> dict(loss, loss_pers, loss_pano, denoise,
        pano_denoise, t, noise, pano_noise)."""
        b, m = latents.shape[:2]
        dev = latents.device
        if t is None:
            t = torch.randint(0, self.num_train_timesteps, (b,), device=dev, generator=generator).long()
        if (noise is None) != (pano_noise is None):
            raise ValueError("noise and pano_noise come from ONE shared field (init_noise): pass both or neither")
        if noise is None:
            pano_noise, noise = init_noise(b, *pano_latent.shape[-2:], *latents.shape[-2:], cameras, dev, generator=generator)
        noise_z = self.add_noise(latents, noise, t)
        pano_noise_z = self.add_noise(pano_latent, pano_noise, t)
        tm = t[:, None].repeat(1, m)
        denoise, pano_denoise = self.mv_base_model(noise_z, pano_noise_z, tm, pers_prompt_embd, pano_prompt_embd, cameras,
                                                   images_layout_cond, pano_layout_cond)
        loss_pers = ops.mse_loss(denoise.float().contiguous(), noise.float().contiguous())
        loss_pano = ops.mse_loss(pano_denoise.float().contiguous(), pano_noise.float().contiguous())
        return dict(loss=loss_pers + loss_pano, loss_pers=loss_pers, loss_pano=loss_pano, denoise=denoise,
                    pano_denoise=pano_denoise, t=t, noise=noise, pano_noise=pano_noise)

    def training_step(self, *args, **kwargs):
        raise NotImplementedError("the backward of the denoiser (UNets, EPPA, LoRA), the optimizer and the gradient "
                                  "all-reduce are not built (SURVEY.md 8f rank 4); TrainingStep.loss is the forward half")
