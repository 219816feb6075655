"""Drop-in ``PLMSSampler`` / ``DDIMSampler`` (reference ``ldm/models/diffusion/plms.py:11-248``,
``ldm/models/diffusion/ddim.py:22-283``): same constructor, ``make_schedule`` and ``sample`` signatures, same return
values, same error conventions — with the per-step tensor work moved into two CUDA kernels behind the C ABI:

* ``pbe_build_unet_input``  — cat(x, z_inpaint, mask) + CFG batch doubling (plms.py:185-186,225; ddim.py:200,209),
* ``pbe_sampler_step``      — CFG combine + Adams-Bashforth extrapolation + pred_x0 / x_prev (plms.py:188-246).

Per-step coefficients are host floats computed once in ``make_schedule`` (the reference's per-step
``torch.full(..., cuda_tensor[index])`` forces three device->host syncs per step).
When the model's U-Net is a ``pbe_b200.UNetModel`` the CFG context is folded once per ``sample()`` and the U-Net is
called directly; any other model goes through its own ``apply_model``.  Both test-bench key spellings are accepted
(``images_inpaint``/``images_mask`` and ``inpaint_image``/``inpaint_mask``) and ``uc`` is broadcast
to the batch (#5).  There is no CPU path: tensors must live on a CUDA device.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .unet import UNetModel


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True):
    """reference ldm/modules/diffusionmodules/util.py:46-60."""
    if ddim_discr_method == "uniform":
        c = num_ddpm_timesteps // num_ddim_timesteps
        ddim_timesteps = np.asarray(list(range(0, num_ddpm_timesteps, c)))
    elif ddim_discr_method == "quad":
        ddim_timesteps = ((np.linspace(0, np.sqrt(num_ddpm_timesteps * .8), num_ddim_timesteps)) ** 2).astype(int)
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    steps_out = ddim_timesteps + 1
    if verbose:
        print(f"Selected timesteps for ddim sampler: {steps_out}")
    return steps_out


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=True):
    """reference util.py:63-74 (alphacums: CPU fp32 tensor)."""
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    if verbose:
        print(f"Selected alphas for ddim sampler: a_t: {alphas}; a_(t-1): {alphas_prev}")
        print(f"For the chosen value of eta, which is {eta}, this results in the following sigma_t schedule for ddim "
              f"sampler {sigmas}")
    return sigmas, alphas, alphas_prev


def _inpaint_kwargs(kwargs):
    """(z_inpaint, mask) from test_model_kwargs under either spelling, or from DDIM's rest=[B,5,h,w]."""
    if "test_model_kwargs" in kwargs:
        k = kwargs["test_model_kwargs"]
        if "images_inpaint" in k:
            return k["images_inpaint"], k["images_mask"]
        if "inpaint_image" in k:
            return k["inpaint_image"], k["inpaint_mask"]
        raise KeyError("images_inpaint")
    if "rest" in kwargs:
        r = kwargs["rest"]
        return r[:, :4], r[:, 4:5]
    return None


class _SamplerBase:
    def __init__(self, model, schedule="linear", match_reference_rng=False, **kwargs):
        super().__init__()
        self.model = model
        # The reference draws noise_like(x.shape) on every x_prev computation even when sigma_t == 0 (plms.py:214 -- twice on
        # the first PLMS step -- and ddim.py:238); the mirror only draws noise it uses.  With match_reference_rng=True the
        # unused draws are made (and discarded) at the same points, so a seeded multi-batch run consumes the CUDA generator
        # exactly like the reference and later internally drawn x_T match.
        self.match_reference_rng = bool(match_reference_rng)
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule

    def register_buffer(self, name, attr):
        if isinstance(attr, torch.Tensor) and attr.device != self.model.device:
            attr = attr.to(self.model.device)
        setattr(self, name, attr)

    # ---- plms.py:24-55 / ddim.py:34-68 -----------------------------------------------------------------------
    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        self.ddim_timesteps = make_ddim_timesteps(ddim_discr_method=ddim_discretize, num_ddim_timesteps=ddim_num_steps,
                                                  num_ddpm_timesteps=self.ddpm_num_timesteps, verbose=verbose)
        alphas_cumprod = self.model.alphas_cumprod
        assert alphas_cumprod.shape[0] == self.ddpm_num_timesteps, "alphas have to be defined for each timestep"
        to_torch = lambda x: x.clone().detach().to(self.model.device, torch.float32)
        ac_cpu = alphas_cumprod.detach().to("cpu", torch.float32)
        self.register_buffer("betas", to_torch(self.model.betas))
        self.register_buffer("alphas_cumprod", to_torch(alphas_cumprod))
        self.register_buffer("alphas_cumprod_prev", to_torch(self.model.alphas_cumprod_prev))
        self.register_buffer("sqrt_alphas_cumprod", to_torch(np.sqrt(ac_cpu)))
        self.register_buffer("sqrt_one_minus_alphas_cumprod", to_torch(np.sqrt(1. - ac_cpu)))
        self.register_buffer("log_one_minus_alphas_cumprod", to_torch(np.log(1. - ac_cpu)))
        self.register_buffer("sqrt_recip_alphas_cumprod", to_torch(np.sqrt(1. / ac_cpu)))
        self.register_buffer("sqrt_recipm1_alphas_cumprod", to_torch(np.sqrt(1. / ac_cpu - 1)))
        ddim_sigmas, ddim_alphas, ddim_alphas_prev = make_ddim_sampling_parameters(
            alphacums=ac_cpu, ddim_timesteps=self.ddim_timesteps, eta=ddim_eta, verbose=verbose)
        self.register_buffer("ddim_sigmas", ddim_sigmas)
        self.register_buffer("ddim_alphas", ddim_alphas)
        self.register_buffer("ddim_alphas_prev", ddim_alphas_prev)
        self.register_buffer("ddim_sqrt_one_minus_alphas", np.sqrt(1. - ddim_alphas))
        acp = self.model.alphas_cumprod_prev.detach().to("cpu", torch.float32)
        sig_orig = ddim_eta * torch.sqrt((1 - acp) / (1 - ac_cpu) * (1 - ac_cpu / acp))
        self.register_buffer("ddim_sigmas_for_original_num_steps", sig_orig)
        # Host-side per-step coefficients, rounded exactly as the reference's per-step torch.full(...) fp32 fills
        # (plms.py:204-207): no device->host sync inside the loop.
        # use_original_steps=True (ddim.py:224-227): the 1000-step tables of the model itself
        self._coef_orig = dict(
            a_t=[float(v) for v in ac_cpu],
            a_prev=[float(v) for v in acp],
            sigma=[float(v) for v in sig_orig.to(torch.float32)],
            sqrt_one_minus_at=[float(v) for v in torch.as_tensor(np.sqrt(1. - ac_cpu)).to(torch.float32)],
        )
        a_cpu = ddim_alphas.detach().to("cpu", torch.float32)
        self._coef = dict(
            a_t=[float(v) for v in a_cpu],
            a_prev=[float(torch.tensor(v, dtype=torch.float32)) for v in ddim_alphas_prev],
            sigma=[float(torch.tensor(float(v), dtype=torch.float32)) for v in torch.as_tensor(np.asarray(ddim_sigmas))],
            sqrt_one_minus_at=[float(v) for v in torch.as_tensor(np.sqrt(1. - a_cpu)).to(torch.float32)],
        )

    # ---- helpers ---------------------------------------------------------------------------------------------
    def _fast_unet(self):
        inner = getattr(getattr(self.model, "model", None), "diffusion_model", None)
        return inner if isinstance(inner, UNetModel) else None

    @staticmethod
    def _require_cuda(t: torch.Tensor):
        if t.device.type != "cuda":
            raise RuntimeError("pbe_b200 samplers run on CUDA (sm_100a) only: no CPU fallback exists")

    def _model_eps(self, unet, x9_in, t_in, c_in, eps_out):
        """U-Net evaluation on the (possibly CFG-doubled) batch.  On the fast path under CFG the input holds the B shared
        samples once (see _build_input / _dup) and the engine evaluates the pair (UNetModel.run_cfg_pair)."""
        if unet is not None:
            # the folded context is state of the shared engine: if anything (an img_callback calling apply_model, a second
            # sampler on the same model) has set another one since this loop set its own, put ours back first
            if getattr(self, "_ctx_mine", None) is not None and unet.context_version != self._ctx_mine[0]:
                unet.set_context(self._ctx_mine[1])
                self._ctx_mine = (unet.context_version, self._ctx_mine[1])
            if eps_out.shape[0] == 2 * x9_in.shape[0]:
                return unet.run_cfg_pair(x9_in, t_in[:x9_in.shape[0]], out=eps_out)
            return unet.run(x9_in, t_in, out=eps_out)
        return self.model.apply_model(x9_in, t_in, c_in).to(torch.float32).contiguous()

    def _step_kernel(self, eps, B, cfg, scale, order, hist, x, index, noise, e_out, x_prev, pred_x0, temperature=1.0,
                     coef=None):
        lib = _lib.load()
        n = x.numel()
        dup = 2 if cfg else 1
        if tuple(eps.shape) != (dup * B,) + tuple(x.shape[1:]) or eps.dtype != torch.float32 or not eps.is_contiguous():
            # the kernel reads n elements per half: anything else would be an out-of-bounds device read
            raise RuntimeError(f"model output has shape {tuple(eps.shape)} ({eps.dtype}), expected "
                               f"{(dup * B,) + tuple(x.shape[1:])} contiguous float32")
        e_uc = eps[:B] if cfg else eps
        e_c = eps[B:] if cfg else None
        h = [(t.data_ptr() if t is not None else None) for t in hist] + [None] * (3 - len(hist))
        c = self._coef if coef is None else coef
        st = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            _lib.check(lib.pbe_sampler_step(
                e_uc.data_ptr(), None if e_c is None else e_c.data_ptr(), float(scale), int(cfg), int(order),
                h[0], h[1], h[2], x.data_ptr(), c["a_t"][index], c["a_prev"][index], c["sigma"][index],
                c["sqrt_one_minus_at"][index], None if noise is None else noise.data_ptr(), float(temperature),
                None if e_out is None else e_out.data_ptr(), x_prev.data_ptr(),
                None if pred_x0 is None else pred_x0.data_ptr(), n, st), "pbe_sampler_step")

    def _corrected_eps(self, eps, B, cfg, scale, score_corrector, corrector_kwargs, x9, t, cond):
        """get_model_output's tail (plms.py:188-194, ddim.py:213-218): CFG combine, then
        ``score_corrector.modify_score(model, e_t, x, t, c, **corrector_kwargs)``.  The corrector is user code, so this runs as
        torch ops and the fused kernel is then called without CFG on the result."""
        assert getattr(self.model, "parameterization", "eps") == "eps"
        e_t = (eps[:B] + scale * (eps[B:] - eps[:B])) if cfg else eps
        e_t = score_corrector.modify_score(self.model, e_t, x9[:B], t[:B], cond, **(corrector_kwargs or {}))
        return e_t.to(torch.float32).contiguous()

    def _build_input(self, x, z, mask, out, dup):
        lib = _lib.load()
        B, _, H, W = x.shape
        st = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            _lib.check(lib.pbe_build_unet_input(x.data_ptr(), z.data_ptr(), mask.data_ptr(), out.data_ptr(), B,
                                                x.shape[1], z.shape[1], mask.shape[1], H * W, dup, st),
                       "pbe_build_unet_input")

    @staticmethod
    def _check_inputs(img, z, m, shape):
        """What torch.cat((x, images_inpaint, images_mask), dim=1) (plms.py:222-225, ddim.py:198-200) would refuse, refused
        here: the kernels take raw pointers and B / H / W from x only."""
        if img.dim() != 4 or tuple(img.shape) != tuple(shape):
            raise RuntimeError(f"x_T has shape {tuple(img.shape)}, expected {tuple(shape)} (batch_size, C, H, W)")
        for k, t in ((1, z), (2, m)):
            if t.dim() != 4:
                raise RuntimeError(f"Tensors must have same number of dimensions: got 4 and {t.dim()}")
            for dim in (0, 2, 3):
                if t.shape[dim] != img.shape[dim]:
                    raise RuntimeError(f"Sizes of tensors must match except in dimension 1. Expected size "
                                       f"{img.shape[dim]} but got size {t.shape[dim]} for tensor number {k} in the list.")

    def _setup(self, cond, shape, x_T, unconditional_guidance_scale, unconditional_conditioning, kwargs):
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T
        self._require_cuda(img)
        img = img.to(torch.float32).contiguous()
        pair = _inpaint_kwargs(kwargs)
        cfg = not (unconditional_conditioning is None or unconditional_guidance_scale == 1.)
        if isinstance(cond, dict):
            cond = torch.cat(cond["c_crossattn"], 1) if "c_crossattn" in cond else cond[list(cond.keys())[0]]
        elif isinstance(cond, (list, tuple)):
            cond = torch.cat(list(cond), 1)
        c_in = cond
        if cfg:
            uc = unconditional_conditioning
            if uc.shape[0] != cond.shape[0]:   # learnable_vector is [1,1,768]
                uc = uc.expand(cond.shape[0], *uc.shape[1:])
            c_in = torch.cat((uc, cond))
        return device, b, img, pair, cfg, c_in


class PLMSSampler(_SamplerBase):
    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        if ddim_eta != 0:
            raise ValueError("ddim_eta must be 0 for PLMS")
        super().make_schedule(ddim_num_steps, ddim_discretize, ddim_eta, verbose)

    @torch.inference_mode()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0., score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, **kwargs):
        if conditioning is not None:
            if isinstance(conditioning, dict):
                cbs = conditioning[list(conditioning.keys())[0]].shape[0]
                if cbs != batch_size:
                    print(f"Warning: Got {cbs} conditionings but batch-size is {batch_size}")
            else:
                if conditioning.shape[0] != batch_size:
                    print(f"Warning: Got {conditioning.shape[0]} conditionings but batch-size is {batch_size}")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        size = (batch_size, C, H, W)
        return self.plms_sampling(conditioning, size, callback=callback, img_callback=img_callback,
                                  quantize_denoised=quantize_x0, mask=mask, x0=x0, ddim_use_original_steps=False,
                                  noise_dropout=noise_dropout, temperature=temperature,
                                  score_corrector=score_corrector, corrector_kwargs=corrector_kwargs, x_T=x_T,
                                  log_every_t=log_every_t, unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning, **kwargs)

    @torch.inference_mode()
    def plms_sampling(self, cond, shape, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None,
                      quantize_denoised=False, mask=None, x0=None, img_callback=None, log_every_t=100,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, **kwargs):
        if ddim_use_original_steps:
            # plms.py:199 reads self.model.ddim_sigmas_for_original_num_steps, which no model has: the reference raises too
            raise AttributeError("'LatentDiffusion' object has no attribute 'ddim_sigmas_for_original_num_steps' "
                                 "(plms.py:199: use_original_steps is not usable with PLMSSampler)")
        if quantize_denoised:
            raise NotImplementedError("quantize_denoised needs a VQ first stage (first_stage_model.quantize, plms.py:211); "
                                      "Paint-by-Example's first stage is AutoencoderKL")
        cond_user = cond
        device, b, img, pair, cfg, c_in = self._setup(cond, shape, x_T, unconditional_guidance_scale,
                                                     unconditional_conditioning, kwargs)
        if pair is None:
            raise KeyError("test_model_kwargs")   # plms.py:220 reads kwargs['test_model_kwargs'] unconditionally
        z_inp, m_inp = (t.to(device=device, dtype=torch.float32).contiguous() for t in pair)

        if "_decode_timesteps" in kwargs:          # DDIMSampler.decode: an explicit prefix of ddim_timesteps (ddim.py:266-267)
            timesteps = kwargs["_decode_timesteps"]
        elif timesteps is None:
            timesteps = self.ddim_timesteps
        else:
            subset_end = int(min(timesteps / self.ddim_timesteps.shape[0], 1) * self.ddim_timesteps.shape[0]) - 1
            timesteps = self.ddim_timesteps[:subset_end]
        intermediates = {"x_inter": [img], "pred_x0": [img]}
        time_range = np.flip(timesteps)
        total_steps = timesteps.shape[0]

        unet = self._fast_unet()
        dup = 2 if cfg else 1
        self._check_inputs(img, z_inp, m_inp, shape)
        _, C, H, W = img.shape
        self._ctx_mine = None
        if unet is not None:
            unet.set_context(c_in)
            self._ctx_mine = (unet.context_version, c_in)
        in_dup = 1 if unet is not None else dup     # the engine takes the CFG pair's shared input once
        x9 = torch.empty((in_dup * b, C + z_inp.shape[1] + m_inp.shape[1], H, W), device=device, dtype=torch.float32)
        eps_buf = torch.empty((dup * b, C, H, W), device=device, dtype=torch.float32)
        ts_dev = torch.as_tensor(np.ascontiguousarray(time_range), device=device, dtype=torch.int64)
        ts_all = ts_dev[:, None].expand(total_steps, dup * b).contiguous()   # one row per step, no per-step H2D
        # step buffers, allocated once per call: two latents (ping-pong), four eps history slots, one pred_x0.  Tensors
        # handed out (intermediates, img_callback) are clones, so reuse is invisible to the caller.
        xbuf = [torch.empty_like(img), torch.empty_like(img)]
        ebuf = [torch.empty_like(img) for _ in range(4)]
        x0buf = torch.empty_like(img)
        old_eps = []
        rng_match = getattr(self, "match_reference_rng", False)

        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            if mask is not None:
                assert x0 is not None
                ts = ts_all[i, :b]
                img_orig = self.model.q_sample(x0, ts)
                img = (img_orig * mask + (1 - mask) * img).contiguous()
            self._build_input(img, z_inp, m_inp, x9, in_dup)
            eps = self._model_eps(unet, x9, ts_all[i], c_in, eps_buf)
            kcfg = cfg
            if score_corrector is not None:      # plms.py:191-193: user code between the CFG combine and the multistep update
                eps = self._corrected_eps(eps, b, cfg, unconditional_guidance_scale, score_corrector, corrector_kwargs, x9,
                                          ts_all[i], cond_user)
                kcfg = False
            x_prev = xbuf[i & 1]
            pred_x0 = torch.empty_like(img) if img_callback else x0buf
            e_t = ebuf[i & 3]
            if len(old_eps) == 0:
                # Pseudo Improved Euler (plms.py:230-235): provisional x_prev from e_t, second U-Net call at t_next
                if rng_match:
                    torch.randn(img.shape, device=device)      # the reference's unused noise_like() draw (plms.py:214)
                self._step_kernel(eps, b, kcfg, unconditional_guidance_scale, 0, [], img, index, None, e_t, x_prev, None)
                self._build_input(x_prev, z_inp, m_inp, x9, in_dup)
                t_nx = ts_all[min(i + 1, total_steps - 1)]
                eps2 = self._model_eps(unet, x9, t_nx, c_in, eps_buf)
                if score_corrector is not None:
                    eps2 = self._corrected_eps(eps2, b, cfg, unconditional_guidance_scale, score_corrector, corrector_kwargs, x9,
                                               t_nx, cond_user)
                self._step_kernel(eps2, b, kcfg, unconditional_guidance_scale, 4, [e_t], img, index, None, None,
                                  x_prev, pred_x0)
            else:
                order = min(len(old_eps), 3)
                hist = list(reversed(old_eps))[:order]   # h1 = most recent
                self._step_kernel(eps, b, kcfg, unconditional_guidance_scale, order, hist, img, index, None, e_t,
                                  x_prev, pred_x0)
            if rng_match:
                torch.randn(img.shape, device=device)
            img = x_prev
            old_eps.append(e_t)
            if len(old_eps) >= 4:
                old_eps.pop(0)
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates["x_inter"].append(img.clone() if index else img)
                intermediates["pred_x0"].append(pred_x0 if img_callback else pred_x0.clone())
        return img, intermediates


class DDIMSampler(_SamplerBase):
    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0., score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, disable_tqdm=False, **kwargs):
        if conditioning is not None:
            if isinstance(conditioning, dict):
                cbs = conditioning[next(iter(conditioning.keys()))].size(0)
                if cbs != batch_size:
                    raise ValueError(f"Warning: Got {cbs} conditionings but batch-size is {batch_size}")
            else:
                if conditioning.size(0) != batch_size:
                    raise ValueError(f"Warning: Got {conditioning.shape[0]} conditionings but batch-size is {batch_size}")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape[-3:]
        size = (batch_size, C, H, W)
        return self.ddim_sampling(conditioning, size, callback=callback, img_callback=img_callback,
                                  quantize_denoised=quantize_x0, mask=mask, x0=x0, ddim_use_original_steps=False,
                                  noise_dropout=noise_dropout, temperature=temperature,
                                  score_corrector=score_corrector, corrector_kwargs=corrector_kwargs, x_T=x_T,
                                  log_every_t=log_every_t, unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning, disable_tqdm=disable_tqdm,
                                  **kwargs)

    @torch.inference_mode()
    def ddim_sampling(self, cond, shape, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None,
                      quantize_denoised=False, mask=None, x0=None, img_callback=None, log_every_t=100,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, disable_tqdm=False, **kwargs):
        if quantize_denoised:
            raise NotImplementedError("quantize_denoised needs a VQ first stage (first_stage_model.quantize, ddim.py:234); "
                                      "Paint-by-Example's first stage is AutoencoderKL")
        cond_user = cond
        device, b, img, pair, cfg, c_in = self._setup(cond, shape, x_T, unconditional_guidance_scale,
                                                     unconditional_conditioning, kwargs)
        if pair is None:
            raise Exception("kwargs must contain either 'test_model_kwargs' or 'rest' key")   # ddim.py:203-204
        z_inp, m_inp = (t.to(device=device, dtype=torch.float32).contiguous() for t in pair)
        coef = None
        if "_decode_timesteps" in kwargs:          # DDIMSampler.decode: an explicit prefix of the timesteps (ddim.py:266-267)
            timesteps = kwargs["_decode_timesteps"]
            coef = self._coef_orig if ddim_use_original_steps else None
        elif ddim_use_original_steps:              # ddim.py:152-153,159-160: all ddpm_num_timesteps steps, the model's own tables
            timesteps = np.arange(self.ddpm_num_timesteps)
            coef = self._coef_orig
        elif timesteps is None:
            timesteps = self.ddim_timesteps
        else:
            subset_end = int(min(timesteps / self.ddim_timesteps.shape[0], 1) * self.ddim_timesteps.shape[0]) - 1
            timesteps = self.ddim_timesteps[:subset_end]
        intermediates = {"x_inter": [img], "pred_x0": [img]}
        time_range = np.flip(timesteps)
        total_steps = timesteps.shape[0]
        sig_tab = (coef or self._coef)["sigma"]

        unet = self._fast_unet()
        dup = 2 if cfg else 1
        self._check_inputs(img, z_inp, m_inp, shape)
        _, C, H, W = img.shape
        self._ctx_mine = None
        if unet is not None:
            unet.set_context(c_in)
            self._ctx_mine = (unet.context_version, c_in)
        in_dup = 1 if unet is not None else dup     # the engine takes the CFG pair's shared input once
        x9 = torch.empty((in_dup * b, C + z_inp.shape[1] + m_inp.shape[1], H, W), device=device, dtype=torch.float32)
        eps_buf = torch.empty((dup * b, C, H, W), device=device, dtype=torch.float32)
        ts_dev = torch.as_tensor(np.ascontiguousarray(time_range), device=device, dtype=torch.int64)
        ts_all = ts_dev[:, None].expand(total_steps, dup * b).contiguous()
        xbuf = [torch.empty_like(img), torch.empty_like(img)]   # per-call step buffers (see plms_sampling)
        x0buf = torch.empty_like(img)
        rng_match = getattr(self, "match_reference_rng", False)

        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            if mask is not None:
                assert x0 is not None
                img_orig = self.model.q_sample(x0, ts_all[i, :b])
                img = (img_orig * mask + (1. - mask) * img).contiguous()
            self._build_input(img, z_inp, m_inp, x9, in_dup)
            eps = self._model_eps(unet, x9, ts_all[i], c_in, eps_buf)
            kcfg = cfg
            if score_corrector is not None:      # ddim.py:216-218
                eps = self._corrected_eps(eps, b, cfg, unconditional_guidance_scale, score_corrector, corrector_kwargs, x9,
                                          ts_all[i], cond_user)
                kcfg = False
            noise = None
            if sig_tab[index] != 0.0:
                # ddim.py:238: noise = sigma_t * noise_like(...) * temperature -- both products happen in the kernel, in that order
                noise = torch.randn(img.shape, device=device)
                if noise_dropout > 0.:
                    noise = torch.nn.functional.dropout(noise, p=noise_dropout)
                noise = noise.contiguous()
            elif rng_match:
                torch.randn(img.shape, device=device)          # the reference draws even when sigma_t == 0
            x_prev = xbuf[i & 1]
            pred_x0 = torch.empty_like(img) if img_callback else x0buf
            self._step_kernel(eps, b, kcfg, unconditional_guidance_scale, 0, [], img, index, noise, None, x_prev, pred_x0,
                              temperature=temperature, coef=coef)
            img = x_prev
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates["x_inter"].append(img.clone() if index else img)
                intermediates["pred_x0"].append(pred_x0 if img_callback else pred_x0.clone())
        return img, intermediates

    @torch.no_grad()
    def decode(self, x_latent, cond, t_start, unconditional_guidance_scale=1.0, unconditional_conditioning=None,
               use_original_steps=False, disable_tqdm=False, **kwargs):
        """ddim.py:262-283: run the last `t_start` DDIM steps from `x_latent` (after make_schedule / a sample() call).
        In this fork p_sample_ddim insists on the inpainting inputs (ddim.py:198-204) and decode() does not forward any,
        so the reference's decode always raises; the same exception is raised here when no `test_model_kwargs` / `rest`
        is given, and -- as an extension -- the steps run when one is."""
        if _inpaint_kwargs(kwargs) is None:
            raise Exception("kwargs must contain either 'test_model_kwargs' or 'rest' key")   # ddim.py:203-204
        timesteps = (np.arange(self.ddpm_num_timesteps) if use_original_steps else self.ddim_timesteps)[:t_start]
        x_dec, _ = self.ddim_sampling(cond, tuple(x_latent.shape), x_T=x_latent, ddim_use_original_steps=use_original_steps,
                                      unconditional_guidance_scale=unconditional_guidance_scale,
                                      unconditional_conditioning=unconditional_conditioning, disable_tqdm=disable_tqdm,
                                      _decode_timesteps=timesteps, **kwargs)
        return x_dec

    @torch.no_grad()
    def stochastic_encode(self, x0, t, use_original_steps=False, noise=None):
        """ddim.py:245-258."""
        if use_original_steps:
            sac, s1m = self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod
        else:
            sac = torch.sqrt(self.ddim_alphas)
            s1m = self.ddim_sqrt_one_minus_alphas
        if noise is None:
            noise = torch.randn_like(x0)
        shape = (x0.shape[0],) + (1,) * (x0.dim() - 1)
        return sac.to(x0.device).gather(-1, t).reshape(shape) * x0 + s1m.to(x0.device).gather(-1, t).reshape(shape) * noise
