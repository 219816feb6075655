"""The oracle (oracle/*.py) is pinned two ways: against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py; runs anywhere) and, where /root/reference is mounted, against the live reference."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import reference_bridge as R
from oracle import sampler_ref as S
from oracle import unet_ref as U


def _golden(golden_dir, name):
    idx = json.load(open(os.path.join(golden_dir, "golden_index.json")))
    a = np.load(os.path.join(golden_dir, name + ".npy"))
    assert hashlib.sha256(a.astype(np.float32).tobytes()).hexdigest() == idx[name]["sha256"], "golden file corrupted"
    return torch.from_numpy(a), idx[name]


@pytest.fixture(scope="module")
def small():
    cfg = U.SMALL_CFG
    sd = U.make_state_dict(cfg, 321)
    req = S.synthetic_request(2, 32, 32, seed=321)
    return cfg, sd, req


def _cfg_inputs(req, B):
    x9 = torch.cat((req["x_T"], req["z_inpaint"], req["mask"]), 1)
    return torch.cat([x9] * 2), torch.cat((req["uc"].expand(B, 1, 768), req["c"]))


def test_state_dict_keys_match_reference(golden_dir, small):
    cfg, sd, _ = small
    idx = json.load(open(os.path.join(golden_dir, "golden_index.json")))
    assert sorted(sd.keys()) == idx["state_dict_keys"]["keys"]
    assert len(U.param_shapes(U.V1_CFG)) == 686  # 686 tensors
    assert sum(int(np.prod(s)) for s in U.param_shapes(U.V1_CFG).values()) == 859_535_364


@pytest.mark.parametrize("tval", [981, 1])
def test_unet_oracle_matches_reference_golden(golden_dir, small, tval):
    cfg, sd, req = small
    x_in, c_in = _cfg_inputs(req, 2)
    e = U.unet_forward(sd, cfg, x_in, torch.full((4,), tval, dtype=torch.int64), c_in)
    g, _ = _golden(golden_dir, f"small_unet_eps_t{tval}")
    assert torch.allclose(e, g, atol=2e-5, rtol=0), (e - g).abs().max()


def test_plms_oracle_matches_reference_golden(golden_dir, small):
    cfg, sd, req = small
    om = S.OracleModel(sd, cfg)
    out = S.plms_sample(om, 8, req["x_T"], req["c"], req["uc"], 5.0, req["z_inpaint"], req["mask"])
    g, _ = _golden(golden_dir, "small_plms8_final")
    assert om.calls == 9  # S + 1: the first step evaluates twice (plms.py:230-235)
    assert torch.allclose(out, g, atol=5e-4, rtol=0), (out - g).abs().max()


def test_ddim_oracle_matches_reference_golden(golden_dir, small):
    cfg, sd, req = small
    om = S.OracleModel(sd, cfg)
    out = S.ddim_sample(om, 5, req["x_T"], req["c"], req["uc"], 5.0, req["z_inpaint"], req["mask"])
    g, _ = _golden(golden_dir, "small_ddim5_final")
    assert om.calls == 5
    assert torch.allclose(out, g, atol=5e-4, rtol=0), (out - g).abs().max()


def test_schedule_anchors():
    """SURVEY.md Appendix A anchors, computed there with the reference's own functions."""
    buf = S.make_schedule_buffers()
    ac = buf["alphas_cumprod"]
    assert abs(ac[0].item() - 0.99915) < 1e-6 and abs(ac[999].item() - 0.0046601) < 1e-7
    tab = S.ddim_tables(ac, 50)
    assert list(tab["timesteps"][:3]) == [1, 21, 41] and tab["timesteps"][-1] == 981
    assert abs(tab["alphas"][49].item() - 0.00577550009) < 1e-9
    assert abs(tab["alphas_prev"][49].item() - 0.00728172716) < 1e-9
    assert abs(tab["alphas"][0].item() - 0.99829602) < 1e-7
    tab20 = S.ddim_tables(ac, 20)
    assert list(tab20["timesteps"][:2]) == [1, 51] and tab20["timesteps"][-1] == 951
    assert abs(tab20["alphas"][19].item() - 0.00815500412) < 1e-9


@pytest.mark.skipif(not R.available(), reason="/root/reference not mounted (GPU box)")
def test_oracle_bit_equals_live_reference(small):
    cfg, sd, req = small
    ref = R.build_reference_unet(cfg, sd)
    x_in, c_in = _cfg_inputs(req, 2)
    t = torch.full((4,), 501, dtype=torch.int64)
    with torch.no_grad():
        e_ref = ref(x_in, t, context=c_in)
    e = U.unet_forward(sd, cfg, x_in, t, c_in)
    assert torch.equal(e, e_ref)
    model = R.StubLatentDiffusion(ref)
    smp = R.reference_sampler("plms", model)
    out_ref, _ = smp.sample(S=4, conditioning=req["c"], batch_size=2, shape=[4, 32, 32], verbose=False,
                            unconditional_guidance_scale=5.0, unconditional_conditioning=req["uc"].expand(2, 1, 768),
                            eta=0.0, x_T=req["x_T"],
                            test_model_kwargs=dict(images_inpaint=req["z_inpaint"], images_mask=req["mask"]))
    out = S.plms_sample(S.OracleModel(sd, cfg), 4, req["x_T"], req["c"], req["uc"], 5.0, req["z_inpaint"], req["mask"])
    assert torch.equal(out, out_ref)


@pytest.mark.skipif(not R.available(), reason="/root/reference not mounted (GPU box)")
@pytest.mark.parametrize("case", ["eta", "eta_temperature", "mask_x0", "plms_mask_x0"])
def test_oracle_stochastic_paths_bit_equal_live_reference(small, case):
    """The sampler paths that draw noise -- DDIM eta > 0 (ddim.py:229-241), temperature != 1 (the (sigma * randn) * T order),
    and the mask / x0 blend of both samplers (ddim.py:168-171, plms.py:150-153) -- consume torch's generator at the same
    points and in the same order as the unmodified reference: same seed, bit-identical latents."""
    cfg, sd, req = small
    model = R.StubLatentDiffusion(R.build_reference_unet(cfg, sd))
    om = S.OracleModel(sd, cfg)
    g = torch.Generator().manual_seed(99)
    x0 = torch.randn(2, 4, 32, 32, generator=g)
    bm = (torch.rand(2, 1, 32, 32, generator=g) > 0.5).float()
    kw = dict(conditioning=req["c"], batch_size=2, shape=[4, 32, 32], verbose=False, unconditional_guidance_scale=5.0,
              unconditional_conditioning=req["uc"].expand(2, 1, 768), x_T=req["x_T"],
              test_model_kwargs=dict(images_inpaint=req["z_inpaint"], images_mask=req["mask"]))
    args = (req["x_T"], req["c"], req["uc"], 5.0, req["z_inpaint"], req["mask"])
    torch.manual_seed(1234)
    if case == "eta":
        out_ref, _ = R.reference_sampler("ddim", model).sample(S=4, eta=0.7, disable_tqdm=True, **kw)
        torch.manual_seed(1234)
        out = S.ddim_sample(om, 4, *args, eta=0.7)
    elif case == "eta_temperature":
        out_ref, _ = R.reference_sampler("ddim", model).sample(S=4, eta=0.5, temperature=0.8, disable_tqdm=True, **kw)
        torch.manual_seed(1234)
        out = S.ddim_sample(om, 4, *args, eta=0.5, temperature=0.8)
    elif case == "mask_x0":
        out_ref, _ = R.reference_sampler("ddim", model).sample(S=4, eta=0.0, mask=bm, x0=x0, disable_tqdm=True, **kw)
        torch.manual_seed(1234)
        out = S.ddim_sample(om, 4, *args, blend_mask=bm, x0=x0)
    else:
        out_ref, _ = R.reference_sampler("plms", model).sample(S=4, eta=0.0, mask=bm, x0=x0, **kw)
        torch.manual_seed(1234)
        out = S.plms_sample(om, 4, *args, blend_mask=bm, x0=x0)
    assert torch.equal(out, out_ref)


class ToyCorrector:
    """A score corrector in the reference's calling convention (plms.py:191-193): elementwise fp32, so CPU == GPU bit for bit."""

    def modify_score(self, model, e_t, x, t, c, gain=1.0):
        return e_t * 0.875 + x[:, :4] * 0.0625 * gain


@pytest.mark.skipif(not R.available(), reason="/root/reference not mounted (GPU box)")
@pytest.mark.parametrize("kind", ["plms", "ddim"])
def test_oracle_score_corrector_bit_equals_live_reference(small, kind):
    cfg, sd, req = small
    model = R.StubLatentDiffusion(R.build_reference_unet(cfg, sd))
    kw = dict(S=4, conditioning=req["c"], batch_size=2, shape=[4, 32, 32], verbose=False, unconditional_guidance_scale=5.0,
              unconditional_conditioning=req["uc"].expand(2, 1, 768), x_T=req["x_T"], eta=0.0,
              score_corrector=ToyCorrector(), corrector_kwargs=dict(gain=2.0),
              test_model_kwargs=dict(images_inpaint=req["z_inpaint"], images_mask=req["mask"]))
    if kind == "ddim":
        kw["disable_tqdm"] = True
    out_ref, _ = R.reference_sampler(kind, model).sample(**kw)
    fn = S.plms_sample if kind == "plms" else S.ddim_sample
    out = fn(S.OracleModel(sd, cfg), 4, req["x_T"], req["c"], req["uc"], 5.0, req["z_inpaint"], req["mask"],
             score_corrector=ToyCorrector(), corrector_kwargs=dict(gain=2.0))
    assert torch.equal(out, out_ref)


@pytest.mark.skipif(not R.available(), reason="/root/reference not mounted (GPU box)")
def test_oracle_stochastic_encode_bit_equals_live_reference(small):
    cfg, sd, req = small
    smp = R.reference_sampler("ddim", R.StubLatentDiffusion(R.build_reference_unet(cfg, sd)))
    smp.make_schedule(ddim_num_steps=10, ddim_eta=0.0, verbose=False)
    g = torch.Generator().manual_seed(5)
    x0, noise = torch.randn(2, 4, 32, 32, generator=g), torch.randn(2, 4, 32, 32, generator=g)
    t = torch.tensor([7, 2])
    ref = smp.stochastic_encode(x0, t, noise=noise)
    out = S.stochastic_encode(S.OracleModel(sd, cfg), 10, x0, t, noise)
    assert torch.equal(out, ref)


# ---- VAE decode: oracle/vae_ref.py ------------------------------------------------------------
@pytest.mark.parametrize("tag", ["small", "v1"])
def test_vae_oracle_matches_reference_golden(golden_dir, tag):
    """oracle.vae_ref.decode == the reference Decoder(post_quant_conv(z)) golden (same seeds), bit for bit on CPU fp32."""
    from oracle import vae_ref as V
    g, meta = _golden(golden_dir, f"{tag}_vae_decode")
    cfg = V.SMALL_VAE_CFG if tag == "small" else V.V1_VAE_CFG
    sd = V.make_state_dict(cfg, meta["weight_seed"])
    z = V.synthetic_latents(meta["B"], meta["hw"], meta["hw"], seed=meta["latent_seed"])
    with torch.no_grad():
        img = V.decode(sd, cfg, z)
    assert img.shape == g.shape
    assert (img - g).abs().max().item() <= 2e-5 * g.abs().max().item()


def test_vae_state_dict_keys_match_reference(golden_dir):
    from oracle import vae_ref as V
    idx = json.load(open(os.path.join(golden_dir, "golden_index.json")))
    assert sorted(V.param_shapes(V.V1_VAE_CFG).keys()) == idx["vae_state_dict_keys"]["keys"]


def test_vae_oracle_bit_equals_live_reference():
    if not R.available():
        pytest.skip("/root/reference not mounted")
    from oracle import vae_ref as V
    cfg = V.SMALL_VAE_CFG
    sd = V.make_state_dict(cfg, 11)
    dec = R.build_reference_vae_decode(cfg, sd)
    z = V.synthetic_latents(2, 8, 16, seed=5)
    with torch.no_grad():
        assert torch.equal(dec(z), V.decode(sd, cfg, z))


@pytest.mark.parametrize("tag", ["small", "v1"])
def test_vae_encode_oracle_matches_reference_golden(golden_dir, tag):
    """oracle.vae_ref.encode_moments == the reference quant_conv(Encoder(x)) golden (same seeds) on CPU fp32."""
    from oracle import vae_ref as V
    g, meta = _golden(golden_dir, f"{tag}_vae_encode_moments")
    cfg = V.SMALL_VAE_CFG if tag == "small" else V.V1_VAE_CFG
    sd = V.make_state_dict(cfg, meta["weight_seed"])
    x = V.synthetic_images(meta["B"], meta["hw"], meta["hw"], seed=meta["image_seed"])
    with torch.no_grad():
        mom = V.encode_moments(sd, cfg, x)
    assert mom.shape == g.shape
    assert (mom - g).abs().max().item() <= 2e-5 * g.abs().max().item()


def test_vae_encode_oracle_bit_equals_live_reference():
    if not R.available():
        pytest.skip("/root/reference not mounted")
    from oracle import vae_ref as V
    cfg = V.SMALL_VAE_CFG
    sd = V.make_state_dict(cfg, 11)
    enc = R.build_reference_vae_encode(cfg, sd)
    x = V.synthetic_images(2, 32, 48, seed=5)
    with torch.no_grad():
        assert torch.equal(enc(x), V.encode_moments(sd, cfg, x))


# ---- conditioning front-end: oracle/clip_ref.py ------------------------------------------------
@pytest.mark.parametrize("tag", ["small", "v1"])
def test_clip_oracle_matches_reference_golden(golden_dir, tag):
    """oracle.clip_ref.encode == live transformers CLIPVisionModel -> reference xf mapper -> final_ln golden."""
    from oracle import clip_ref as K
    g, meta = _golden(golden_dir, f"{tag}_clip_embed")
    cfg = K.SMALL_CLIP_CFG if tag == "small" else K.V1_CLIP_CFG
    sd = K.make_state_dict(cfg, meta["weight_seed"])
    x = K.synthetic_exemplars(meta["B"], cfg["image_size"], seed=meta["image_seed"])
    with torch.no_grad():
        z = K.encode(sd, cfg, x)
    assert z.shape == g.shape
    assert (z - g).abs().max().item() <= 1e-4 * g.abs().max().item()


def test_clip_state_dict_keys_match_reference(golden_dir):
    from oracle import clip_ref as K
    idx = json.load(open(os.path.join(golden_dir, "golden_index.json")))
    assert sorted(K.param_shapes(K.V1_CLIP_CFG).keys()) == idx["clip_state_dict_keys"]["keys"]


def test_clip_oracle_matches_live_reference():
    if not R.available():
        pytest.skip("/root/reference not mounted")
    pytest.importorskip("transformers")
    from oracle import clip_ref as K
    cfg = K.SMALL_CLIP_CFG
    sd = K.make_state_dict(cfg, 11)
    enc = R.build_reference_clip_embedder(cfg, sd)
    x = K.synthetic_exemplars(3, cfg["image_size"], seed=5)
    with torch.no_grad():
        a, b = enc(x), K.encode(sd, cfg, x)
    assert (a - b).abs().max().item() <= 1e-5 * a.abs().max().item()


# ------------------------------------------------------------------------------------------------------------------
# pre-processing oracle (oracle/preprocess_ref.py) against the LIVE torchvision transforms / torch interpolate the reference
# scripts call (scripts/inference.py:106-124, 305-332).  These are library functions present on every box, not /root/reference.
def test_preprocess_oracle_matches_live_torchvision_transforms():
    import numpy as np
    import torchvision
    from PIL import Image
    from oracle import preprocess_ref as P
    img_u8, mask_u8 = P.synthetic_u8_request(2, 48, 40, seed=5)
    for mean, std in ((P.HALF, P.HALF), (P.CLIP_MEAN, P.CLIP_STD)):
        tf = torchvision.transforms.Compose([torchvision.transforms.ToTensor(), torchvision.transforms.Normalize(mean, std)])
        ref = torch.stack([tf(Image.fromarray(img_u8[b].numpy())) for b in range(2)])
        assert torch.equal(P.normalize_u8(img_u8, mean, std), ref)
    # scripts/inference.py:311-318, restated literally on the PIL objects
    for b in range(2):
        image_tensor = torchvision.transforms.Compose([torchvision.transforms.ToTensor(), torchvision.transforms.Normalize(
            (0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])(Image.fromarray(img_u8[b].numpy())).unsqueeze(0)
        mask = np.array(Image.fromarray(mask_u8[b].numpy()).convert("L"))[None, None]
        mask = 1 - mask.astype(np.float32) / 255.0
        mask[mask < 0.5] = 0
        mask[mask >= 0.5] = 1
        mask_tensor = torch.from_numpy(mask)
        image, m, inpaint = P.prepare_inpaint(img_u8[b:b + 1], mask_u8[b:b + 1], binarize=True)
        assert torch.equal(image, image_tensor) and torch.equal(m, mask_tensor) and torch.equal(inpaint, image_tensor * mask_tensor)
        # ldm/data/test_bench_dataset.py:89-98
        mask_tensor = 1 - torchvision.transforms.ToTensor()(Image.fromarray(mask_u8[b].numpy()).convert("L"))
        image, m, inpaint = P.prepare_inpaint(img_u8[b:b + 1], mask_u8[b:b + 1], binarize=False)
        assert torch.equal(m[0], mask_tensor) and torch.equal(inpaint[0], image_tensor[0] * mask_tensor)
        assert ((m > 0) & (m < 1)).any()           # the soft border survives without the threshold


@pytest.mark.parametrize("H,W,h,w", [(512, 512, 64, 64), (768, 768, 96, 96), (96, 80, 12, 10), (100, 60, 13, 7), (30, 30, 45, 50)])
@pytest.mark.parametrize("antialias", [False, True])
def test_resize_oracle_matches_live_interpolate(H, W, h, w, antialias):
    """Resize([h, w]) on a tensor = F.interpolate(bilinear, align_corners=False[, antialias]); ATen may contract to FMA, so
    general inputs agree to an ulp and {0, 1} masks at the reference's power-of-two factors agree exactly."""
    import torch.nn.functional as F
    import torchvision
    from oracle import preprocess_ref as P
    g = torch.Generator().manual_seed(H + w)
    x = torch.rand(2, 1, H, W, generator=g)
    ref = F.interpolate(x, size=(h, w), mode="bilinear", align_corners=False, antialias=antialias)
    assert (P.resize_bilinear(x, (h, w), antialias) - ref).abs().max().item() <= 5e-7
    assert torch.equal(torchvision.transforms.Resize([h, w], antialias=antialias)(x), ref)
    m = (x > 0.6).float()
    m[:, :, H // 4: H // 2, W // 3:] = 1.0
    ref = F.interpolate(m, size=(h, w), mode="bilinear", align_corners=False, antialias=antialias)
    out = P.resize_bilinear(m, (h, w), antialias)
    if not antialias and H % h == 0 and W % w == 0:
        assert torch.equal(out, ref)
    assert (out - ref).abs().max().item() <= 5e-7
