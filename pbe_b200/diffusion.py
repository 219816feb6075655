"""Host-side mirror of the reference diffusion wrapper around the U-Net, restricted to what the sampling hot path uses.

* ``DiffusionWrapper``  — reference ``ldm/models/diffusion/ddpm.py:468-515`` (conditioning_key='crossattn').
* ``LatentDiffusion``   — reference ``ldm/models/diffusion/latent_diffusion.py:85`` / ``ddpm.py:87``: schedule buffers
  (``register_schedule`` ddpm.py:175-228), ``apply_model`` (latent_diffusion.py:646-743), ``q_sample``,
  ``learnable_vector`` / ``proj_out`` (latent_diffusion.py:111-112).  VAE and CLIP stages are outside the hot path
  (SURVEY.md §8f): they can be attached as ordinary torch modules (``first_stage_model`` / ``cond_stage_model``).

State-dict keys match the reference checkpoint layout (``model.diffusion_model.*``, ``learnable_vector``, ``proj_out.*``).
"""
from __future__ import annotations

import contextlib

import numpy as np
import torch
import torch.nn as nn

from .unet import UNetModel


class DiffusionWrapper(nn.Module):
    def __init__(self, unet_params: dict, conditioning_key: str = "crossattn"):
        super().__init__()
        if conditioning_key != "crossattn":
            raise NotImplementedError("Paint-by-Example uses conditioning_key='crossattn' (configs/v1.yaml:15)")
        self.diffusion_model = UNetModel(**unet_params)
        self.conditioning_key = conditioning_key

    def forward(self, x, t, c_concat=None, c_crossattn=None):
        cc = torch.cat(c_crossattn, 1)
        return self.diffusion_model(x, t, context=cc)


class LatentDiffusion(nn.Module):
    def __init__(self, unet_config=None, timesteps=1000, linear_start=0.00085, linear_end=0.0120,
                 scale_factor=0.18215, channels=4, image_size=64, conditioning_key="crossattn",
                 parameterization="eps", **ignored):
        super().__init__()
        params = dict(unet_config.get("params", unet_config)) if unet_config is not None else {}
        self.model = DiffusionWrapper(params, conditioning_key)
        self.parameterization = parameterization
        self.channels = channels
        self.image_size = image_size
        self.scale_factor = scale_factor
        self.learnable_vector = nn.Parameter(torch.randn((1, 1, 768)), requires_grad=False)
        self.proj_out = nn.Linear(1024, 768)
        self.first_stage_model = None
        self.cond_stage_model = None
        self.register_schedule(timesteps, linear_start, linear_end)

    # ---- ddpm.py:175-228 --------------------------------------------------------------------------------------
    def register_schedule(self, timesteps=1000, linear_start=0.00085, linear_end=0.0120):
        betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64) ** 2).numpy()
        alphas = 1. - betas
        alphas_cumprod = np.cumprod(alphas, axis=0)
        alphas_cumprod_prev = np.append(1., alphas_cumprod[:-1])
        self.num_timesteps = int(timesteps)
        self.linear_start, self.linear_end = linear_start, linear_end
        f32 = lambda a: torch.tensor(a, dtype=torch.float32)
        self.register_buffer("betas", f32(betas))
        self.register_buffer("alphas_cumprod", f32(alphas_cumprod))
        self.register_buffer("alphas_cumprod_prev", f32(alphas_cumprod_prev))
        self.register_buffer("sqrt_alphas_cumprod", f32(np.sqrt(alphas_cumprod)))
        self.register_buffer("sqrt_one_minus_alphas_cumprod", f32(np.sqrt(1. - alphas_cumprod)))
        self.register_buffer("log_one_minus_alphas_cumprod", f32(np.log(1. - alphas_cumprod)))
        self.register_buffer("sqrt_recip_alphas_cumprod", f32(np.sqrt(1. / alphas_cumprod)))
        self.register_buffer("sqrt_recipm1_alphas_cumprod", f32(np.sqrt(1. / alphas_cumprod - 1)))

    @property
    def device(self):
        return self.betas.device

    @contextlib.contextmanager
    def ema_scope(self, context=None):  # use_ema: False (configs/v1.yaml:19) -> no-op, ddpm.py:230-243
        yield None

    # ---- latent_diffusion.py:646-743 (the fold/unfold branch is dead: split_input_params is never set) ---------
    def apply_model(self, x_noisy, t, cond, return_ids=False):
        if isinstance(cond, dict):
            pass
        else:
            if not isinstance(cond, list):
                cond = [cond]
            cond = {"c_crossattn": cond}
        x_recon = self.model(x_noisy, t, **cond)
        if isinstance(x_recon, tuple) and not return_ids:
            return x_recon[0]
        return x_recon

    # ---- ddpm.py q_sample --------------------------------------------------------------------------------------
    def q_sample(self, x_start, t, noise=None):
        noise = torch.randn_like(x_start) if noise is None else noise
        shape = (x_start.shape[0],) + (1,) * (x_start.dim() - 1)
        a = self.sqrt_alphas_cumprod.gather(-1, t).reshape(shape)
        b = self.sqrt_one_minus_alphas_cumprod.gather(-1, t).reshape(shape)
        return a * x_start + b * noise

    # ---- stages outside the hot path ---------------------------------------------------------------------------
    def _need(self, what):
        raise NotImplementedError(f"{what} is outside the accelerated hot path (SURVEY.md §8f); attach the reference "
                                  f"module as `first_stage_model` / `cond_stage_model` to use it")

    def get_learned_conditioning(self, c):
        if self.cond_stage_model is None:
            self._need("CLIP exemplar encoder")
        return self.cond_stage_model(c)

    def encode_first_stage(self, x):
        if self.first_stage_model is None:
            self._need("VAE encoder")
        return self.first_stage_model.encode(x)

    def get_first_stage_encoding(self, encoder_posterior):
        z = encoder_posterior.sample() if hasattr(encoder_posterior, "sample") else encoder_posterior
        return self.scale_factor * z

    def decode_first_stage(self, z, **kw):
        if self.first_stage_model is None:
            self._need("VAE decoder")
        z = 1. / self.scale_factor * z
        return self.first_stage_model.decode(z)
