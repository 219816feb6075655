"""Sampler paths the round-1 suite did not reach (VERDICT r1 "untested paths"): DDIM eta > 0 (the sigma_t * noise term),
temperature != 1 (order of the two products), the mask / x0 blend of both samplers, stochastic_encode, the shape checks in
front of the raw-pointer kernels, and parity at the BASELINE configurations C4 (DDIM-20, batch 1) and C5 (96x96 latent,
batch 4 -> CFG batch 8).

Two kinds of checks:
* bit-exact: the sampler loop + K11 kernel against the oracle (oracle/sampler_ref.py, pinned torch.equal to the live
  reference for exactly these paths in tests/test_oracle_pinned.py) with an elementwise fp32 eps model whose arithmetic is
  identical on CPU and GPU -- every rounding of the update, the order of the generator draws and the buffer reuse are visible;
* tolerance (BASELINE.json): the real bf16 U-Net against the fp32 oracle / the reference's golden trajectory, PSNR >= 40 dB,
  per-call eps rel-L2 <= 1e-2.
"""
import ctypes
import hashlib
import json
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

PSNR_DB = 40.0
EPS_REL_L2 = 1e-2


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def _psnr(a, ref):
    mse = ((a.float() - ref.float()) ** 2).mean().item()
    peak = ref.abs().max().item()
    return 10 * math.log10(peak * peak / max(mse, 1e-30))


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def lib(dev):
    from pbe_b200 import _lib
    return _lib.load()


def _stream():
    return torch.cuda.current_stream().cuda_stream


# ---- K11 with the noise term ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("temperature", [1.0, 0.8])
@pytest.mark.parametrize("cfg", [0, 1])
def test_sampler_step_with_noise_bit_exact_vs_oracle(lib, dev, cfg, temperature):
    """sigma_t != 0: x_prev = sqrt(a_prev) * pred_x0 + dir_xt + (sigma_t * noise) * temperature, rounded in the reference's
    order (ddim.py:238,241; plms.py:214,217).  eta = 0.7 makes sigma_t ~ 0.1."""
    from oracle import sampler_ref as S
    g = torch.Generator().manual_seed(7 + cfg)
    shape = (3, 4, 64, 64)
    eu, ec, x, nz = (torch.randn(shape, generator=g) for _ in range(4))
    tab = S.ddim_tables(S.make_schedule_buffers()["alphas_cumprod"], 20, eta=0.7)
    index = 11
    a_t, a_prev, sig, s1m = (float(tab[k][index]) for k in ("alphas", "alphas_prev", "sigmas", "sqrt_one_minus_alphas"))
    assert sig > 1e-3
    scale = 5.0
    e = eu + scale * (ec - eu) if cfg else eu
    xp_ref, x0_ref = S._x_prev_and_pred_x0(x, e, a_t, a_prev, sig, s1m, nz, temperature)
    d = lambda t: t.to(dev).contiguous()
    eu_d, ec_d, x_d, nz_d = map(d, (eu, ec, x, nz))
    xp, x0 = (torch.empty(shape, device=dev) for _ in range(2))
    f = ctypes.c_float
    rc = lib.pbe_sampler_step(eu_d.data_ptr(), ec_d.data_ptr() if cfg else None, f(scale), cfg, 0, None, None, None,
                              x_d.data_ptr(), f(a_t), f(a_prev), f(sig), f(s1m), nz_d.data_ptr(), f(temperature), None,
                              xp.data_ptr(), x0.data_ptr(), x.numel(), _stream())
    assert rc == 0, lib.pbe_last_error().decode()
    torch.cuda.synchronize()
    assert torch.equal(x0.cpu(), x0_ref)
    assert torch.equal(xp.cpu(), xp_ref)
    # sigma_t != 0 without a noise tensor is refused, not read from NULL
    rc = lib.pbe_sampler_step(eu_d.data_ptr(), ec_d.data_ptr() if cfg else None, f(scale), cfg, 0, None, None, None,
                              x_d.data_ptr(), f(a_t), f(a_prev), f(sig), f(s1m), None, f(temperature), None, xp.data_ptr(),
                              x0.data_ptr(), x.numel(), _stream())
    assert rc != 0


def test_build_unet_input_channel_counts(lib, dev):
    """cat((x, z, mask), 1) for channel splits other than 4 + 4 + 1 (the kernel used to hard-code them)."""
    x, z, m = torch.randn(3, 4, 16, 24, device=dev), torch.randn(3, 5, 16, 24, device=dev), torch.rand(3, 2, 16, 24, device=dev)
    out = torch.empty(6, 11, 16, 24, device=dev)
    rc = lib.pbe_build_unet_input(x.data_ptr(), z.data_ptr(), m.data_ptr(), out.data_ptr(), 3, 4, 5, 2, 16 * 24, 2, _stream())
    assert rc == 0, lib.pbe_last_error().decode()
    torch.cuda.synchronize()
    assert torch.equal(out, torch.cat([torch.cat((x, z, m), 1)] * 2))
    assert lib.pbe_build_unet_input(x.data_ptr(), None, m.data_ptr(), out.data_ptr(), 3, 4, 5, 2, 16 * 24, 2, _stream()) != 0


# ---- whole sampler loops, bit-exact, with an elementwise eps model ------------------------------------------------------
def _toy_eps(x9, t, c):
    """Elementwise fp32 (one IEEE rounding per op, no fusion in eager torch): the same bits on CPU and GPU."""
    b = x9.shape[0]
    tt = (t.to(torch.float32) * 0.001).view(b, 1, 1, 1)
    cc = (c[:, 0, :4] * 0.1).view(b, 4, 1, 1)
    return x9[:, :4] * 0.25 + x9[:, 4:8] * 0.5 - x9[:, 8:9] * 0.125 + tt + cc


class _ToyProduct:
    """What the samplers read from LatentDiffusion, around the toy eps model (generic-model path of the samplers)."""

    def __init__(self, dev):
        from pbe_b200.diffusion import LatentDiffusion
        from oracle import unet_ref as U
        self.inner = LatentDiffusion(unet_config=dict(params=dict(U.SMALL_CFG))).to(dev)
        self.num_timesteps = self.inner.num_timesteps
        self.betas, self.alphas_cumprod = self.inner.betas, self.inner.alphas_cumprod
        self.alphas_cumprod_prev = self.inner.alphas_cumprod_prev
        self.device = self.inner.device
        self.q_sample = self.inner.q_sample

    def apply_model(self, x, t, c):
        return _toy_eps(x, t, c)


class _ToyOracle:
    def __init__(self):
        from oracle import sampler_ref as S
        for k, v in S.make_schedule_buffers().items():
            setattr(self, k, v)
        self.num_timesteps = 1000

    def apply_model(self, x, t, c):
        return _toy_eps(x, t, c)


@pytest.fixture(scope="module")
def toy(dev):
    from oracle import sampler_ref as S
    req = S.synthetic_request(3, 16, 24, seed=77)
    g = torch.Generator().manual_seed(78)
    x0 = torch.randn(3, 4, 16, 24, generator=g)
    bm = (torch.rand(3, 1, 16, 24, generator=g) > 0.4).float()
    return _ToyProduct(dev), _ToyOracle(), req, x0, bm


def _kw(req, dev, B):
    d = lambda t: t.to(dev)
    return dict(conditioning=d(req["c"]), batch_size=B, shape=[4, 16, 24], verbose=False, unconditional_guidance_scale=5.0,
                unconditional_conditioning=d(req["uc"]), x_T=d(req["x_T"]),
                test_model_kwargs=dict(images_inpaint=d(req["z_inpaint"]), images_mask=d(req["mask"])))


@pytest.mark.parametrize("eta,temperature,oracle_on", [(0.5, 1.0, "cpu"), (0.7, 0.8, "cpu"), (0.9, 1.3, "cpu"), (1.0, 1.3, "cuda")])
def test_ddim_eta_and_temperature_bit_exact(toy, dev, eta, temperature, oracle_on):
    """DDIM with eta > 0 and a temperature: the noise drawn from the CUDA generator at the reference's point (after the model
    call, ddim.py:238) enters as (sigma_t * noise) * temperature; 10 steps, every intermediate compared.
    eta = 1.0 runs the oracle's torch ops on the GPU: at index 5 of that schedule (1 - a_prev - sigma_t**2).sqrt() is
    0x1.17d90b01p-1 before rounding, torch's CPU sqrt returns ...90a, the correctly rounded value -- what CUDA's sqrtf, and so
    the reference on a GPU, and this library's host sqrtf give -- is ...90c.  The oracle code is the same either way."""
    from oracle import sampler_ref as S
    from pbe_b200.samplers import DDIMSampler
    prod, orc, req, _, _ = toy
    torch.manual_seed(2024)
    out, inter = DDIMSampler(prod).sample(S=10, eta=eta, temperature=temperature, log_every_t=1, **_kw(req, dev, 3))
    torch.manual_seed(2024)
    rec = []
    o = (lambda t: t.to(dev)) if oracle_on == "cuda" else (lambda t: t)
    ref = S.ddim_sample(orc, 10, o(req["x_T"]), o(req["c"]), o(req["uc"]), 5.0, o(req["z_inpaint"]), o(req["mask"]), record=rec,
                        eta=eta, temperature=temperature, rng_device=dev)
    assert torch.equal(out.cpu(), ref.cpu())
    assert len(inter["x_inter"]) == 11 and len(inter["pred_x0"]) == 11
    for k, r in enumerate(rec):      # buffers are reused between steps: the logged tensors must be snapshots
        assert torch.equal(inter["x_inter"][k + 1].cpu(), r["x_prev"].cpu()), k
        assert torch.equal(inter["pred_x0"][k + 1].cpu(), r["pred_x0"].cpu()), k


@pytest.mark.parametrize("kind", ["ddim", "plms"])
def test_mask_x0_blend_bit_exact(toy, dev, kind):
    """mask / x0 (ddim.py:168-171, plms.py:150-153): img = q_sample(x0, ts) * mask + (1 - mask) * img before every step.
    With match_reference_rng=True the CUDA generator is consumed exactly as the reference consumes it (q_sample's draw plus
    the unused noise_like() draws), so the whole trajectory is bit-identical to the oracle's."""
    from oracle import sampler_ref as S
    from pbe_b200.samplers import DDIMSampler, PLMSSampler
    prod, orc, req, x0, bm = toy
    args = (req["x_T"], req["c"], req["uc"], 5.0, req["z_inpaint"], req["mask"])
    smp = (DDIMSampler if kind == "ddim" else PLMSSampler)(prod, match_reference_rng=True)
    torch.manual_seed(31)
    out, _ = smp.sample(S=10, eta=0.0, mask=bm.to(dev), x0=x0.to(dev), **_kw(req, dev, 3))
    torch.manual_seed(31)
    fn = S.ddim_sample if kind == "ddim" else S.plms_sample
    ref = fn(orc, 10, *args, blend_mask=bm, x0=x0, rng_device=dev)
    assert torch.equal(out.cpu(), ref)
    # without the option the latents differ from the reference's from the second blend on (fewer draws), by design
    torch.manual_seed(31)
    out2, _ = (DDIMSampler if kind == "ddim" else PLMSSampler)(prod).sample(S=10, eta=0.0, mask=bm.to(dev), x0=x0.to(dev),
                                                                          **_kw(req, dev, 3))
    assert not torch.equal(out2, out)


def test_plms_plain_loop_and_callbacks_bit_exact(toy, dev):
    """PLMS-8 (first-step double evaluation, AB2..AB4) with the reused step buffers: final latent, every logged
    intermediate and the tensors handed to img_callback equal the oracle's."""
    from oracle import sampler_ref as S
    from pbe_b200.samplers import PLMSSampler
    prod, orc, req, _, _ = toy
    rec = []
    ref = S.plms_sample(orc, 8, req["x_T"], req["c"], req["uc"], 5.0, req["z_inpaint"], req["mask"], record=rec)
    seen, steps = [], []
    out, inter = PLMSSampler(prod).sample(S=8, eta=0.0, log_every_t=2, img_callback=lambda p, i: seen.append(p),
                                          callback=steps.append, **_kw(req, dev, 3))
    assert torch.equal(out.cpu(), ref)
    assert steps == list(range(8)) and len(seen) == 8
    for k, r in enumerate(rec):
        assert torch.equal(seen[k].cpu(), r["pred_x0"]), k       # callback tensors are not overwritten by later steps
    logged = [r for r in rec if r["index"] % 2 == 0 or r["index"] == 7]
    assert len(inter["x_inter"]) == 1 + len(logged)
    for a, r in zip(inter["x_inter"][1:], logged):
        assert torch.equal(a.cpu(), r["x_prev"])
    out_b, _ = PLMSSampler(prod).sample(S=8, eta=0.0, **_kw(req, dev, 3))
    assert torch.equal(out_b, out)


@pytest.mark.parametrize("kind", ["plms", "ddim"])
def test_score_corrector_bit_exact(toy, dev, kind):
    """score_corrector.modify_score between the CFG combine and the update (plms.py:191-193, ddim.py:216-218): user code, so the
    combine runs as torch ops and the fused kernel takes the corrected eps without CFG -- bit-identical to the oracle."""
    from oracle import sampler_ref as S
    from pbe_b200.samplers import DDIMSampler, PLMSSampler
    from test_oracle_pinned import ToyCorrector
    prod, orc, req, _, _ = toy
    cls, fn = (PLMSSampler, S.plms_sample) if kind == "plms" else (DDIMSampler, S.ddim_sample)
    out, _ = cls(prod).sample(S=8, eta=0.0, score_corrector=ToyCorrector(), corrector_kwargs=dict(gain=2.0), **_kw(req, dev, 3))
    ref = fn(orc, 8, req["x_T"], req["c"], req["uc"], 5.0, req["z_inpaint"], req["mask"], score_corrector=ToyCorrector(),
             corrector_kwargs=dict(gain=2.0))
    assert torch.equal(out.cpu(), ref)
    plain, _ = cls(prod).sample(S=8, eta=0.0, **_kw(req, dev, 3))
    assert not torch.equal(plain, out)


def test_ddim_original_steps_decode(toy, dev):
    """use_original_steps=True (ddim.py:152-160,224-227,262-267): the model's own 1000-step tables instead of the DDIM
    sub-sequence.  In this fork both samplers read `self.model.ddim_sigmas_for_original_num_steps`, which no model has, so the
    reference raises AttributeError on this path (PLMSSampler here raises the same); DDIMSampler implements what the line
    means (upstream CompVis reads the sampler's own buffer).  Checked against a direct restatement of the update with the
    model's tables: the last 6 of the 1000 steps, eta = 0."""
    from oracle import sampler_ref as S
    from pbe_b200.samplers import DDIMSampler, PLMSSampler
    prod, orc, req, _, _ = toy
    smp = DDIMSampler(prod)
    smp.make_schedule(ddim_num_steps=10, ddim_eta=0.0, verbose=False)
    d = lambda t: t.to(dev)
    kw = dict(unconditional_guidance_scale=5.0, unconditional_conditioning=d(req["uc"]),
              test_model_kwargs=dict(images_inpaint=d(req["z_inpaint"]), images_mask=d(req["mask"])))
    out = smp.decode(d(req["x_T"]), d(req["c"]), 6, use_original_steps=True, **kw)
    # the restatement runs its torch ops on the GPU: torch's CPU sqrt is not correctly rounded for every input (see
    # test_ddim_eta_and_temperature_bit_exact), CUDA's -- what the reference uses on a GPU -- and the library's host sqrtf are
    buf = S.make_schedule_buffers()
    ac, acp = buf["alphas_cumprod"], buf["alphas_cumprod_prev"]
    s1m = torch.tensor(np.sqrt(1. - ac.numpy().astype(np.float32)))
    img = d(req["x_T"])
    for step in range(5, -1, -1):
        t = torch.full((3,), step, dtype=torch.int64, device=dev)
        e = S._cfg_eps(orc, torch.cat((img, d(req["z_inpaint"]), d(req["mask"])), 1), t, d(req["c"]), d(req["uc"]), 5.0)
        img, _ = S._x_prev_and_pred_x0(img, e, ac[step], acp[step], 0.0, s1m[step])
    img = img.cpu()
    assert torch.equal(out.cpu(), img)
    with pytest.raises(AttributeError, match="ddim_sigmas_for_original_num_steps"):
        PLMSSampler(prod).plms_sampling(d(req["c"]), (3, 4, 16, 24), ddim_use_original_steps=True, **kw)


def test_stochastic_encode_bit_exact(toy, dev):
    from oracle import sampler_ref as S
    from pbe_b200.samplers import DDIMSampler
    prod, orc, req, x0, _ = toy
    smp = DDIMSampler(prod)
    smp.make_schedule(ddim_num_steps=10, ddim_eta=0.0, verbose=False)
    g = torch.Generator().manual_seed(3)
    noise = torch.randn(3, 4, 16, 24, generator=g)
    t = torch.tensor([9, 4, 0])
    out = smp.stochastic_encode(x0.to(dev), t.to(dev), noise=noise.to(dev))
    assert torch.equal(out.cpu(), S.stochastic_encode(orc, 10, x0, t, noise))
    torch.manual_seed(8)
    a = smp.stochastic_encode(x0.to(dev), t.to(dev))
    torch.manual_seed(8)
    n2 = torch.randn_like(x0.to(dev))
    assert torch.equal(a.cpu(), S.stochastic_encode(orc, 10, x0, t, n2.cpu()))


# ---- what torch.cat would refuse is refused before a raw pointer reaches a kernel ----------------------------------------
def test_sampler_shape_checks(toy, dev):
    from pbe_b200.samplers import DDIMSampler, PLMSSampler
    prod, orc, req, _, _ = toy
    d = lambda t: t.to(dev)
    for cls in (PLMSSampler, DDIMSampler):
        kw = _kw(req, dev, 3)
        kw["test_model_kwargs"] = dict(images_inpaint=d(req["z_inpaint"][:1]), images_mask=d(req["mask"]))   # batch-1 z
        with pytest.raises(RuntimeError, match="Sizes of tensors must match"):
            cls(prod).sample(S=4, eta=0.0, **kw)
        kw = _kw(req, dev, 3)
        kw["test_model_kwargs"] = dict(images_inpaint=d(req["z_inpaint"]),
                                       images_mask=d(torch.ones(3, 1, 128, 192)))                          # image-resolution mask
        with pytest.raises(RuntimeError, match="Sizes of tensors must match"):
            cls(prod).sample(S=4, eta=0.0, **kw)
        kw = _kw(req, dev, 3)
        kw["x_T"] = d(req["x_T"][:2])                                                                      # x_T batch != batch_size
        with pytest.raises(RuntimeError, match="x_T has shape"):
            cls(prod).sample(S=4, eta=0.0, **kw)

    class Bad(_ToyProduct):
        def apply_model(self, x, t, c):
            return _toy_eps(x, t, c)[:, :, :8]          # wrong spatial size

    with pytest.raises(RuntimeError, match="model output has shape"):
        PLMSSampler(Bad(dev)).sample(S=4, eta=0.0, **_kw(req, dev, 3))


# ---- the real U-Net on the stochastic path ---------------------------------------------------------------------------
def test_small_ddim_eta_vs_oracle(dev):
    """DDIM-8 with eta = 0.5 through the bf16 CUDA U-Net vs the fp32 oracle fed the same CUDA-generator noise."""
    from oracle import sampler_ref as S, unet_ref as U
    from pbe_b200.diffusion import LatentDiffusion
    from pbe_b200.samplers import DDIMSampler
    cfg = U.SMALL_CFG
    sd = U.make_state_dict(cfg, 321)
    req = S.synthetic_request(2, 32, 32, seed=321)
    model = LatentDiffusion(unet_config=dict(params=dict(cfg)))
    model.load_state_dict({"model.diffusion_model." + k: v for k, v in sd.items()}, strict=False)
    model = model.to(dev).eval()
    d = lambda t: t.to(dev)
    torch.manual_seed(5)
    out, _ = DDIMSampler(model).sample(S=8, conditioning=d(req["c"]), batch_size=2, shape=[4, 32, 32], verbose=False,
                                       unconditional_guidance_scale=5.0, unconditional_conditioning=d(req["uc"]), eta=0.5,
                                       x_T=d(req["x_T"]),
                                       test_model_kwargs=dict(images_inpaint=d(req["z_inpaint"]), images_mask=d(req["mask"])))
    torch.manual_seed(5)
    ref = S.ddim_sample(S.OracleModel(sd, cfg), 8, req["x_T"], req["c"], req["uc"], 5.0, req["z_inpaint"], req["mask"],
                        eta=0.5, rng_device=dev)
    p = _psnr(out.cpu(), ref)
    print(f"small DDIM-8 eta=0.5: PSNR vs fp32 oracle = {p:.2f} dB")
    assert p >= PSNR_DB, p


# ---- BASELINE configurations C4 and C5 at config size ------------------------------------------------------------------
@pytest.fixture(scope="module")
def v1(dev):
    from oracle import unet_ref as U
    from pbe_b200.diffusion import LatentDiffusion
    cfg = U.V1_CFG
    sd = U.make_state_dict(cfg, 321)
    model = LatentDiffusion(unet_config=dict(params=dict(cfg)))
    model.load_state_dict({"model.diffusion_model." + k: v for k, v in sd.items()}, strict=False)
    return cfg, sd, model.to(dev).eval()


def test_v1_ddim20_c4_vs_reference_golden(v1, dev, golden_dir):
    """BASELINE config C4 (test.sh:1-9): DDIM 20 steps, batch 1, 64x64 latent, scale 5 -- against the trajectory the
    unmodified reference computed on CPU (tests/golden/make_golden.py v1_ddim20)."""
    from oracle import sampler_ref as S
    from pbe_b200.samplers import DDIMSampler
    idx = json.load(open(os.path.join(golden_dir, "golden_index.json")))
    if "v1_ddim20_final" not in idx:
        pytest.skip("golden v1_ddim20_final not generated")
    a = np.load(os.path.join(golden_dir, "v1_ddim20_final.npy"))
    assert hashlib.sha256(a.astype(np.float32).tobytes()).hexdigest() == idx["v1_ddim20_final"]["sha256"]
    g = torch.from_numpy(a)
    cfg, sd, model = v1
    req = S.synthetic_request(1, 64, 64, seed=321)
    d = lambda t: t.to(dev)
    out, _ = DDIMSampler(model).sample(S=20, conditioning=d(req["c"]), batch_size=1, shape=[4, 64, 64], verbose=False,
                                       unconditional_guidance_scale=5.0, unconditional_conditioning=d(req["uc"]), eta=0.0,
                                       x_T=d(req["x_T"]),
                                       test_model_kwargs=dict(images_inpaint=d(req["z_inpaint"]), images_mask=d(req["mask"])))
    p = _psnr(out.cpu(), g)
    print(f"v1 DDIM-20 C4: PSNR vs reference CPU trajectory = {p:.2f} dB, rel-L2 = {_rel(out.cpu(), g):.3e}")
    assert p >= PSNR_DB, p


def test_v1_c5_96_batch4_eps_and_plms10_vs_oracle_on_gpu(v1, dev):
    """BASELINE config C5 at config size: 96x96 latent, batch 4 (CFG batch 8).  (a) one CFG U-Net call, eps rel-L2 <= 1e-2;
    (b) a PLMS-10 trajectory (11 U-Net calls), PSNR >= 40 dB -- both against the fp32 oracle run with torch on the same GPU
    (TF32 off), two samples at a time (its attention matrix is 2.7 GB per sample at 9216 tokens)."""
    from oracle import sampler_ref as S, unet_ref as U
    from pbe_b200.samplers import PLMSSampler
    cfg, sd, model = v1
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    B, hw = 4, 96
    req = S.synthetic_request(B, hw, hw, seed=96)
    d = lambda t: t.to(dev)
    x9 = torch.cat((req["x_T"], req["z_inpaint"], req["mask"]), 1).to(dev)
    x_in, c_in = torch.cat([x9] * 2), torch.cat((req["uc"].expand(B, 1, 768), req["c"])).to(dev)
    t = torch.full((2 * B,), 741, dtype=torch.int64, device=dev)
    eps = model.apply_model(x_in, t, c_in)
    ref = torch.cat([U.unet_forward(sd_dev, cfg, x_in[i:i + 2], t[i:i + 2], c_in[i:i + 2]) for i in range(0, 2 * B, 2)])
    r = _rel(eps, ref)
    print(f"v1 C5 (96x96, CFG batch 8): eps rel-L2 vs fp32 oracle on GPU = {r:.3e}")
    assert r <= EPS_REL_L2, r

    class ChunkedOracle(S.OracleModel):
        def apply_model(self, x_noisy, t_, cond):
            self.calls += 1
            return torch.cat([U.unet_forward(self.sd, self.cfg, x_noisy[i:i + 2], t_[i:i + 2], cond[i:i + 2])
                              for i in range(0, x_noisy.shape[0], 2)])

    out, _ = PLMSSampler(model).sample(S=10, conditioning=d(req["c"]), batch_size=B, shape=[4, hw, hw], verbose=False,
                                       unconditional_guidance_scale=5.0, unconditional_conditioning=d(req["uc"]), eta=0.0,
                                       x_T=d(req["x_T"]),
                                       test_model_kwargs=dict(images_inpaint=d(req["z_inpaint"]), images_mask=d(req["mask"])))
    om = ChunkedOracle(sd, cfg, device=dev)
    ref_out = S.plms_sample(om, 10, d(req["x_T"]), d(req["c"]), d(req["uc"]), 5.0, d(req["z_inpaint"]), d(req["mask"]))
    assert om.calls == 11
    p = _psnr(out, ref_out)
    print(f"v1 C5 PLMS-10 (96x96, batch 4): PSNR vs fp32 oracle trajectory on GPU = {p:.2f} dB")
    assert p >= PSNR_DB, p
