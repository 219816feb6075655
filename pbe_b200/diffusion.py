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
                 parameterization="eps", first_stage_config=None, cond_stage_config=None, cond_stage_forward=None,
                 first_stage_key="inpaint", cond_stage_key="image", use_ema=False, ckpt_path=None, **ignored):
        """Accepts everything ``configs/v1.yaml:4-73`` passes (training-only keys -- scheduler_config, monitor,
        u_cond_percent, cond_stage_trainable, num_timesteps_cond, log_every_t ... -- land in ``ignored``).  The first / cond
        stages are built from their configs exactly as ``instantiate_first_stage`` / ``instantiate_cond_stage`` do
        (latent_diffusion.py:215-240): through ``instantiate_from_config`` on the ``target:`` the yaml names, which resolves to
        this package's VAE / CLIP mirrors once ``pbe_b200.install()`` has run (or when the yaml names them directly)."""
        super().__init__()
        if use_ema:
            raise NotImplementedError("use_ema=True is a training feature (configs/v1.yaml:19 sets False)")
        if ckpt_path is not None:
            raise NotImplementedError("load weights with load_state_dict (scripts/inference.py:60-65), not ckpt_path")
        if unet_config is not None and "target" in unet_config and not str(unet_config["target"]).endswith(".UNetModel"):
            raise NotImplementedError(f"unet_config.target {unet_config['target']!r}: only UNetModel is accelerated")
        params = dict(unet_config.get("params", unet_config)) if unet_config is not None else {}
        params.pop("target", None)
        self.model = DiffusionWrapper(params, conditioning_key)
        self.parameterization = parameterization
        self.channels = channels
        self.image_size = image_size
        self.scale_factor = scale_factor
        self.first_stage_key, self.cond_stage_key = first_stage_key, cond_stage_key
        self.cond_stage_forward = cond_stage_forward
        self.learnable_vector = nn.Parameter(torch.randn((1, 1, 768)), requires_grad=False)
        self.proj_out = nn.Linear(1024, 768)
        self.first_stage_model = None
        self.cond_stage_model = None
        if first_stage_config is not None:
            from .dropin import instantiate_from_config
            self.first_stage_model = instantiate_from_config(first_stage_config).eval()
        if cond_stage_config is not None and cond_stage_config not in ("__is_first_stage__", "__is_unconditional__"):
            from .dropin import instantiate_from_config
            self.cond_stage_model = instantiate_from_config(cond_stage_config).eval()
        elif cond_stage_config == "__is_first_stage__":
            self.cond_stage_model = self.first_stage_model
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
        """latent_diffusion.py:264-276."""
        m = self.cond_stage_model
        if m is None:
            self._need("CLIP exemplar encoder")
        if self.cond_stage_forward is None:
            if hasattr(m, "encode") and callable(m.encode):
                c = m.encode(c)
                if hasattr(c, "mode") and callable(c.mode) and not isinstance(c, torch.Tensor):
                    c = c.mode()
            else:
                c = m(c)
        else:
            assert hasattr(m, self.cond_stage_forward)
            c = getattr(m, self.cond_stage_forward)(c)
        return c

    def encode_first_stage(self, x):
        if self.first_stage_model is None:
            self._need("VAE encoder")
        return self.first_stage_model.encode(x)

    def get_first_stage_encoding(self, encoder_posterior):
        """latent_diffusion.py:255-262."""
        if isinstance(encoder_posterior, torch.Tensor):
            z = encoder_posterior
        elif hasattr(encoder_posterior, "sample"):
            z = encoder_posterior.sample()
        else:
            raise NotImplementedError(f"encoder_posterior of type '{type(encoder_posterior)}' not yet implemented")
        return self.scale_factor * z

    def decode_first_stage(self, z, **kw):
        if self.first_stage_model is None:
            self._need("VAE decoder")
        z = 1. / self.scale_factor * z
        return self.first_stage_model.decode(z)
