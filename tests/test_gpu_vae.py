"""VAE decode on the GPU, through pbe_b200.AutoencoderKL.decode -> pbe_vae_decode (C ABI), against
the reference-generated goldens and the fp32 oracle.  Tolerance: relative L2 <= 1e-2 and PSNR >= 40 dB (peak = max
|reference image|), the same bf16-operand / fp32-accumulate bar as the denoising path."""
import hashlib
import json
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def _psnr(a, ref):
    mse = ((a.float() - ref.float()) ** 2).mean().item()
    peak = ref.abs().max().item()
    return 10 * math.log10(peak * peak / max(mse, 1e-30))


def _golden(golden_dir, name):
    idx = json.load(open(os.path.join(golden_dir, "golden_index.json")))
    a = np.load(os.path.join(golden_dir, name + ".npy"))
    assert hashlib.sha256(a.astype(np.float32).tobytes()).hexdigest() == idx[name]["sha256"]
    return torch.from_numpy(a), idx[name]


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda:0")


def _make_vae(cfg, sd, dev):
    from pbe_b200.vae import AutoencoderKL
    dd = dict(double_z=True, z_channels=cfg["z_channels"], resolution=256, in_channels=3, out_ch=cfg["out_ch"], ch=cfg["ch"],
              ch_mult=list(cfg["ch_mult"]), num_res_blocks=cfg["num_res_blocks"], attn_resolutions=[], dropout=0.0)
    m = AutoencoderKL(ddconfig=dd, embed_dim=cfg["embed_dim"])
    m.load_state_dict(sd, strict=True)
    return m.to(dev).eval()


@pytest.mark.parametrize("tag", ["small", "v1"])
def test_vae_decode_vs_reference_golden(dev, golden_dir, tag):
    from oracle import vae_ref as V
    g, meta = _golden(golden_dir, f"{tag}_vae_decode")
    cfg = V.SMALL_VAE_CFG if tag == "small" else V.V1_VAE_CFG
    sd = V.make_state_dict(cfg, meta["weight_seed"])
    vae = _make_vae(cfg, sd, dev)
    z = V.synthetic_latents(meta["B"], meta["hw"], meta["hw"], seed=meta["latent_seed"])
    img = vae.decode(z.to(dev)).cpu()
    assert img.shape == g.shape and torch.isfinite(img).all()
    print(f"{tag} VAE decode: rel-L2 = {_rel(img, g):.3e}, PSNR = {_psnr(img, g):.1f} dB")
    assert _rel(img, g) <= 1e-2 and _psnr(img, g) >= 40.0


def test_vae_decode_512_batch_vs_oracle_on_gpu(dev):
    """The real geometry (64x64 latent -> 512x512 image, the 4096-token mid attention) and a ragged batch, against the
    fp32 oracle executed with torch on the same GPU."""
    from oracle import vae_ref as V
    cfg = V.V1_VAE_CFG
    sd = V.make_state_dict(cfg, 321)
    vae = _make_vae(cfg, sd, dev)
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    for B, h, w in ((3, 64, 64), (1, 32, 48)):
        z = V.synthetic_latents(B, h, w, seed=B * 100 + h).to(dev)
        img = vae.decode(z)
        with torch.no_grad():
            ref = torch.cat([V.decode(sd_dev, cfg, z[i:i + 1]) for i in range(B)])
        print(f"v1 VAE decode B={B} {h}x{w}: rel-L2 = {_rel(img, ref):.3e}, PSNR = {_psnr(img, ref):.1f} dB")
        assert img.shape == (B, 3, 8 * h, 8 * w)
        assert _rel(img, ref) <= 1e-2 and _psnr(img, ref) >= 40.0


def test_decode_first_stage_uses_the_vae(dev):
    """LatentDiffusion.decode_first_stage(z) == first_stage_model.decode(z / scale_factor) (latent_diffusion.py:444-508)."""
    from oracle import unet_ref as U, vae_ref as V
    from pbe_b200.diffusion import LatentDiffusion
    cfg = V.SMALL_VAE_CFG
    vae = _make_vae(cfg, V.make_state_dict(cfg, 3), dev)
    ld = LatentDiffusion(unet_config=dict(params=dict(U.SMALL_CFG)))
    ld.first_stage_model = vae
    z = V.synthetic_latents(1, 16, 16, seed=9).to(dev)
    a = ld.decode_first_stage(z * ld.scale_factor)
    b = vae.decode(z)
    assert _rel(a, b) < 1e-3


@pytest.mark.parametrize("tag", ["small", "v1"])
def test_vae_encode_vs_reference_golden(dev, golden_dir, tag):
    from oracle import vae_ref as V
    g, meta = _golden(golden_dir, f"{tag}_vae_encode_moments")
    cfg = V.SMALL_VAE_CFG if tag == "small" else V.V1_VAE_CFG
    vae = _make_vae(cfg, V.make_state_dict(cfg, meta["weight_seed"]), dev)
    x = V.synthetic_images(meta["B"], meta["hw"], meta["hw"], seed=meta["image_seed"])
    post = vae.encode(x.to(dev))
    mom = post.parameters.cpu()
    assert mom.shape == g.shape and torch.isfinite(mom).all()
    print(f"{tag} VAE encode: rel-L2 = {_rel(mom, g):.3e}, PSNR = {_psnr(mom, g):.1f} dB")
    assert _rel(mom, g) <= 1e-2 and _psnr(mom, g) >= 40.0
    # DiagonalGaussianDistribution semantics (distributions.py:24-41)
    mean, logvar = torch.chunk(post.parameters, 2, dim=1)
    assert torch.equal(post.mode(), mean) and torch.equal(post.logvar, logvar.clamp(-30.0, 20.0))
    torch.manual_seed(0)
    s1 = post.sample()
    torch.manual_seed(0)
    noise = torch.randn(mean.shape).to(dev)
    assert torch.equal(s1, mean + post.std * noise)


def test_vae_encode_512_vs_oracle_on_gpu(dev):
    """The real geometry (512x512 image -> 64x64 latent, asymmetric-padding stride-2 convs, 4096-token mid attention)."""
    from oracle import vae_ref as V
    cfg = V.V1_VAE_CFG
    sd = V.make_state_dict(cfg, 321)
    vae = _make_vae(cfg, sd, dev)
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    for B, H, W in ((2, 512, 512), (1, 256, 384)):
        x = V.synthetic_images(B, H, W, seed=B * 10 + H).to(dev)
        mom = vae.encode(x).parameters
        with torch.no_grad():
            ref = torch.cat([V.encode_moments(sd_dev, cfg, x[i:i + 1]) for i in range(B)])
        print(f"v1 VAE encode B={B} {H}x{W}: rel-L2 = {_rel(mom, ref):.3e}, PSNR = {_psnr(mom, ref):.1f} dB")
        assert mom.shape == (B, 8, H // 8, W // 8)
        assert _rel(mom, ref) <= 1e-2 and _psnr(mom, ref) >= 40.0


def test_vae_round_trip_first_stage(dev):
    """encode_first_stage -> get_first_stage_encoding -> decode_first_stage through LatentDiffusion (latent_diffusion.py
    :571-610, :444-508): shapes, scale factor and the posterior plumbing."""
    from oracle import unet_ref as U, vae_ref as V
    from pbe_b200.diffusion import LatentDiffusion
    cfg = V.SMALL_VAE_CFG
    ld = LatentDiffusion(unet_config=dict(params=dict(U.SMALL_CFG)))
    ld.first_stage_model = _make_vae(cfg, V.make_state_dict(cfg, 3), dev)
    x = V.synthetic_images(2, 64, 64, seed=4).to(dev)
    post = ld.encode_first_stage(x)
    z = ld.get_first_stage_encoding(post)
    assert z.shape == (2, 4, 16, 16)
    assert torch.allclose(ld.get_first_stage_encoding(post.mode()), ld.scale_factor * post.mean)
    img = ld.decode_first_stage(z)
    assert img.shape == x.shape and torch.isfinite(img).all()


def test_decode_to_uint8_matches_reference_postprocessing(dev):
    """scripts/inference.py:346-348,379-380: clamp((x+1)/2, 0, 1) -> HWC -> 255 * x -> astype(uint8), bit for bit."""
    import numpy as np
    from oracle import vae_ref as V
    cfg = V.SMALL_VAE_CFG
    vae = _make_vae(cfg, V.make_state_dict(cfg, 3), dev)
    z = V.synthetic_latents(2, 16, 16, seed=2).to(dev) * 2.0       # large enough to exercise both clamps
    u8 = vae.decode_to_uint8(z)
    x = vae.decode(z)
    x = torch.clamp((x + 1.0) / 2.0, min=0.0, max=1.0).cpu().permute(0, 2, 3, 1).numpy()
    ref = (255. * x).astype(np.uint8)
    assert u8.shape == (2, 64, 64, 3) and u8.dtype == torch.uint8
    assert (ref == 0).any() and (ref == 255).any()
    assert np.array_equal(u8.cpu().numpy(), ref)
