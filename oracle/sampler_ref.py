"""ORACLE — test infrastructure only (never imported by the product path ``pbe_b200/``).

Restatement of the reference's schedule and PLMS / DDIM samplers in plain torch fp32, op for op and in the same order
(SURVEY.md Appendix A), so a fused CUDA step can be checked bit-for-bit.  Citations are relative to /root/reference.
"""
from __future__ import annotations

import numpy as np
import torch

from .unet_ref import unet_forward


def make_schedule_buffers(timesteps=1000, linear_start=0.00085, linear_end=0.0120):
    """DDPM.register_schedule, ldm/models/diffusion/ddpm.py:175-197 with make_beta_schedule('linear'),
    ldm/modules/diffusionmodules/util.py:21-25: float64 math, stored as fp32."""
    betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64) ** 2).numpy()
    alphas = 1.0 - betas
    alphas_cumprod = np.cumprod(alphas, axis=0)
    alphas_cumprod_prev = np.append(1.0, alphas_cumprod[:-1])
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    return dict(betas=f32(betas), alphas_cumprod=f32(alphas_cumprod), alphas_cumprod_prev=f32(alphas_cumprod_prev),
                sqrt_alphas_cumprod=f32(np.sqrt(alphas_cumprod)),                     # ddpm.py:200-201: float64 sqrt, then fp32
                sqrt_one_minus_alphas_cumprod=f32(np.sqrt(1. - alphas_cumprod)))


def ddim_tables(alphas_cumprod: torch.Tensor, S: int, eta: float = 0.0, num_ddpm=1000):
    """make_ddim_timesteps ('uniform') + make_ddim_sampling_parameters, util.py:46-74, and the derived
    sqrt_one_minus table of plms.py:48-51.  Returns fp32 torch tensors exactly as the per-step torch.full fills
    see them (plms.py:204-207)."""
    c = num_ddpm // S
    ts = np.asarray(list(range(0, num_ddpm, c))) + 1
    ac = alphas_cumprod.cpu()
    alphas = ac[ts]                                                             # torch fp32
    alphas_prev = np.asarray([ac[0]] + ac[ts[:-1]].tolist())                    # numpy float64 of fp32 values
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    sqrt_one_minus = np.sqrt(1.0 - alphas)                                      # stays a torch fp32 tensor
    return dict(timesteps=ts,
                alphas=alphas.to(torch.float32),
                alphas_prev=torch.tensor(alphas_prev, dtype=torch.float32),
                sigmas=torch.as_tensor(np.asarray(sigmas), dtype=torch.float32),
                sqrt_one_minus_alphas=torch.as_tensor(sqrt_one_minus).to(torch.float32))


class OracleModel:
    """What the samplers need from LatentDiffusion: schedule buffers + apply_model
    (latent_diffusion.py:646-654,739 -> DiffusionWrapper.forward ddpm.py:484-486 -> UNetModel.forward)."""

    def __init__(self, sd, cfg, device="cpu"):
        self.sd = {k: v.to(device) for k, v in sd.items()}
        self.cfg = cfg
        self.device = torch.device(device)
        self.num_timesteps = 1000
        for k, v in make_schedule_buffers().items():
            setattr(self, k, v.to(device))
        self.calls = 0

    def apply_model(self, x_noisy, t, cond):
        if isinstance(cond, dict):
            cond = torch.cat(cond["c_crossattn"], 1)
        elif isinstance(cond, (list, tuple)):
            cond = torch.cat(list(cond), 1)
        self.calls += 1
        return unet_forward(self.sd, self.cfg, x_noisy, t, cond)


def _x_prev_and_pred_x0(x, e, a_t, a_prev, sigma_t, sqrt_one_minus_at, noise=None, temperature=1.):
    """get_x_prev_and_pred_x0, plms.py:202-219 / the tail of p_sample_ddim, ddim.py:221-242.
    noise: the N(0,1) draw of noise_like(); None = the term is dropped (it is sigma_t * randn = +-0 when eta = 0)."""
    b = x.shape[0]
    full = lambda v: torch.full((b, 1, 1, 1), float(v), device=x.device)
    a_t, a_prev, sigma_t, s1m = full(a_t), full(a_prev), full(sigma_t), full(sqrt_one_minus_at)
    pred_x0 = (x - s1m * e) / a_t.sqrt()
    dir_xt = (1. - a_prev - sigma_t ** 2).sqrt() * e
    x_prev = a_prev.sqrt() * pred_x0 + dir_xt
    if noise is not None:
        x_prev = x_prev + sigma_t * noise * temperature        # plms.py:214,217 / ddim.py:238,241
    return x_prev, pred_x0


def q_sample(model, x_start, t, noise):
    """DDPM.q_sample, ddpm.py:348-351 (extract_into_tensor of the two sqrt tables)."""
    shape = (x_start.shape[0],) + (1,) * (x_start.dim() - 1)
    sac, s1m = model.sqrt_alphas_cumprod.to(x_start.device), model.sqrt_one_minus_alphas_cumprod.to(x_start.device)
    return sac.gather(-1, t).reshape(shape) * x_start + s1m.gather(-1, t).reshape(shape) * noise


def _cfg_eps(model, x9, t, c, uc, scale, score_corrector=None, corrector_kwargs=None):
    """get_model_output, plms.py:181-194 (ddim.py:205-218): CFG combine, then the optional score corrector."""
    if uc is None or scale == 1.:
        e_t = model.apply_model(x9, t, c)
    else:
        if uc.shape[0] != c.shape[0]:
            uc = uc.expand(c.shape[0], *uc.shape[1:])
        x_in = torch.cat([x9] * 2)
        t_in = torch.cat([t] * 2)
        c_in = torch.cat((uc, c))
        e_u, e_c = model.apply_model(x_in, t_in, c_in).chunk(2)
        e_t = e_u + scale * (e_c - e_u)
    if score_corrector is not None:
        e_t = score_corrector.modify_score(model, e_t, x9, t, c, **(corrector_kwargs or {}))
    return e_t


@torch.no_grad()
def plms_sample(model, S, x_T, c, uc, scale, z_inpaint, mask, record=None, blend_mask=None, x0=None, rng_device=None,
                score_corrector=None, corrector_kwargs=None):
    """PLMSSampler.plms_sampling + p_sample_plms, plms.py:118-248 (eta = 0).  With blend_mask / x0 (plms.py:150-153) the
    generator of `rng_device` is consumed exactly as the reference does: q_sample's randn_like, then one (unused, sigma_t = 0)
    noise_like() draw per get_x_prev_and_pred_x0 call (plms.py:214) -- two on the first step."""
    tab = ddim_tables(model.alphas_cumprod, S)
    ts = tab["timesteps"]
    time_range = np.flip(ts)
    total = len(ts)
    img = x_T
    b = img.shape[0]
    old_eps = []
    rdev = img.device if rng_device is None else torch.device(rng_device)
    draw = (lambda: torch.randn(img.shape, device=rdev)) if blend_mask is not None else (lambda: None)
    for i, step in enumerate(time_range):
        index = total - i - 1
        t = torch.full((b,), int(step), device=img.device, dtype=torch.int64)
        t_next = torch.full((b,), int(time_range[min(i + 1, total - 1)]), device=img.device, dtype=torch.int64)
        coef = (tab["alphas"][index], tab["alphas_prev"][index], tab["sigmas"][index],
                tab["sqrt_one_minus_alphas"][index])
        if blend_mask is not None:
            img_orig = q_sample(model, x0, t, torch.randn(x0.shape, device=rdev).to(img.device))
            img = img_orig * blend_mask + (1 - blend_mask) * img
        x9 = torch.cat((img, z_inpaint, mask), dim=1)
        e_t = _cfg_eps(model, x9, t, c, uc, scale, score_corrector, corrector_kwargs)
        if len(old_eps) == 0:
            draw()
            x_prev, _ = _x_prev_and_pred_x0(img, e_t, *coef)
            e_next = _cfg_eps(model, torch.cat((x_prev, z_inpaint, mask), dim=1), t_next, c, uc, scale, score_corrector,
                              corrector_kwargs)
            e_prime = (e_t + e_next) / 2
        elif len(old_eps) == 1:
            e_prime = (3 * e_t - old_eps[-1]) / 2
        elif len(old_eps) == 2:
            e_prime = (23 * e_t - 16 * old_eps[-1] + 5 * old_eps[-2]) / 12
        else:
            e_prime = (55 * e_t - 59 * old_eps[-1] + 37 * old_eps[-2] - 9 * old_eps[-3]) / 24
        draw()
        x_prev, pred_x0 = _x_prev_and_pred_x0(img, e_prime, *coef)
        if record is not None:
            record.append(dict(index=index, e_t=e_t.clone(), x_prev=x_prev.clone(), pred_x0=pred_x0.clone()))
        img = x_prev
        old_eps.append(e_t)
        if len(old_eps) >= 4:
            old_eps.pop(0)
    return img


@torch.no_grad()
def ddim_sample(model, S, x_T, c, uc, scale, z_inpaint, mask, record=None, eta=0.0, temperature=1., blend_mask=None,
                x0=None, rng_device=None, score_corrector=None, corrector_kwargs=None):
    """DDIMSampler.ddim_sampling + p_sample_ddim, ddim.py:136-242.  With eta > 0 (or blend_mask / x0, ddim.py:168-171) the
    noise is drawn from torch's global generator of `rng_device` (default: the latent's device) at the reference's points
    and in its order: q_sample's randn_like first, then noise_like() after the model call."""
    tab = ddim_tables(model.alphas_cumprod, S, eta=eta)
    ts = tab["timesteps"]
    time_range = np.flip(ts)
    total = len(ts)
    img = x_T
    b = img.shape[0]
    rdev = img.device if rng_device is None else torch.device(rng_device)
    for i, step in enumerate(time_range):
        index = total - i - 1
        t = torch.full((b,), int(step), device=img.device, dtype=torch.int64)
        if blend_mask is not None:
            img_orig = q_sample(model, x0, t, torch.randn(x0.shape, device=rdev).to(img.device))
            img = img_orig * blend_mask + (1. - blend_mask) * img
        x9 = torch.cat((img, z_inpaint, mask), dim=1)
        e_t = _cfg_eps(model, x9, t, c, uc, scale, score_corrector, corrector_kwargs)
        noise = torch.randn(img.shape, device=rdev).to(img.device) if (eta != 0.0 or blend_mask is not None) else None
        x_prev, pred_x0 = _x_prev_and_pred_x0(img, e_t, tab["alphas"][index], tab["alphas_prev"][index],
                                              tab["sigmas"][index], tab["sqrt_one_minus_alphas"][index], noise, temperature)
        if record is not None:
            record.append(dict(index=index, e_t=e_t.clone(), x_prev=x_prev.clone(), pred_x0=pred_x0.clone()))
        img = x_prev
    return img


def stochastic_encode(model, S, x0, t, noise, eta=0.0):
    """DDIMSampler.stochastic_encode, ddim.py:245-258 (use_original_steps=False): t indexes the DDIM tables."""
    tab = ddim_tables(model.alphas_cumprod, S, eta=eta)
    shape = (x0.shape[0],) + (1,) * (x0.dim() - 1)
    sac = torch.sqrt(tab["alphas"]).to(x0.device)
    s1m = tab["sqrt_one_minus_alphas"].to(x0.device)
    return sac.gather(-1, t).reshape(shape) * x0 + s1m.gather(-1, t).reshape(shape) * noise


def synthetic_request(B, h, w, seed=321, ctx_dim=768, device="cpu"):
    """Synthetic edit requests: x_T, z_inpaint ~ N(0,1), bbox-hole mask (1 = keep, 0 = hole),
    exemplar token c, shared learned unconditional token uc."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    x_T = torch.randn(B, 4, h, w, generator=g)
    z = torch.randn(B, 4, h, w, generator=g)
    mask = torch.ones(B, 1, h, w)
    for b in range(B):
        frac = 0.05 + 0.45 * torch.rand(1, generator=g).item()
        bh = max(1, int(round(h * frac ** 0.5)))
        bw = max(1, int(round(w * frac ** 0.5)))
        y0 = int(torch.randint(0, h - bh + 1, (1,), generator=g).item())
        x0 = int(torch.randint(0, w - bw + 1, (1,), generator=g).item())
        mask[b, :, y0:y0 + bh, x0:x0 + bw] = 0.0
    c = torch.randn(B, 1, ctx_dim, generator=g)
    uc = torch.randn(1, 1, ctx_dim, generator=torch.Generator(device="cpu").manual_seed(4242))
    to = lambda t: t.to(device)
    return dict(x_T=to(x_T), z_inpaint=to(z), mask=to(mask), c=to(c), uc=to(uc))
