"""The whole request path of scripts/inference.py:300-348 on the library's own stages — exemplar -> CLIP front-end ->
proj_out -> conditioning; masked image -> VAE encode -> latent; PLMS under CFG over the U-Net; VAE decode -> uint8 —
against the same pipeline composed from the fp32 oracles (small configurations of all three networks)."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _psnr(a, ref):
    mse = ((a.float() - ref.float()) ** 2).mean().item()
    peak = ref.abs().max().item()
    return 10 * math.log10(peak * peak / max(mse, 1e-30))


def test_inference_script_flow_vs_oracle_pipeline():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda:0")
    from oracle import clip_ref as K, sampler_ref as S, unet_ref as U, vae_ref as V
    from pbe_b200.clip import FrozenCLIPImageEmbedder
    from pbe_b200.diffusion import LatentDiffusion
    from pbe_b200.samplers import PLMSSampler
    from pbe_b200.vae import AutoencoderKL

    B, steps, scale = 2, 8, 5.0
    ucfg, vcfg, kcfg = U.SMALL_CFG, V.SMALL_VAE_CFG, K.SMALL_CLIP_CFG
    usd, vsd, ksd = U.make_state_dict(ucfg, 321), V.make_state_dict(vcfg, 321), K.make_state_dict(kcfg, 321)
    g = torch.Generator().manual_seed(5)
    proj_w = torch.randn(768, kcfg["width"], generator=g) / math.sqrt(kcfg["width"])      # model.proj_out (latent_diffusion.py:112)
    proj_b = 0.05 * torch.randn(768, generator=g)
    uc = torch.randn(1, 1, 768, generator=g)                                              # model.learnable_vector (:111)
    f = 2 ** (len(vcfg["ch_mult"]) - 1)
    H = W = 32 * f                                                                        # image -> 32 x 32 latent
    image = V.synthetic_images(B, H, W, seed=6)
    mask = torch.ones(B, 1, H, W)
    mask[:, :, H // 4: H // 2, W // 4: 3 * W // 4] = 0.0                                  # 1 = keep, 0 = hole
    ref_img = K.synthetic_exemplars(B, kcfg["image_size"], seed=7)
    x_T = torch.randn(B, 4, 32, 32, generator=g)
    sf = 0.18215

    # ---------------- oracle pipeline (fp32, torch on the GPU) ----------------
    d = lambda t: t.to(dev)
    usd_d, vsd_d, ksd_d = ({k: d(v) for k, v in sd.items()} for sd in (usd, vsd, ksd))
    with torch.no_grad():
        c_ref = F.linear(K.encode(ksd_d, kcfg, d(ref_img)), d(proj_w), d(proj_b))
        mom = V.encode_moments(vsd_d, vcfg, d(image * mask))
        z_ref = sf * torch.chunk(mom, 2, dim=1)[0]                                        # posterior.mode()
        m_lat = F.interpolate(d(mask), size=(32, 32), mode="nearest")
        om = S.OracleModel(usd, ucfg, device=dev)
        lat_ref = S.plms_sample(om, steps, d(x_T), c_ref, d(uc).expand(B, 1, 768), scale, z_ref, m_lat)
        img_ref = V.decode(vsd_d, vcfg, lat_ref / sf)

    # ---------------- the library, wired like scripts/inference.py ----------------
    model = LatentDiffusion(unet_config=dict(params=dict(ucfg)))
    model.load_state_dict({"model.diffusion_model." + k: v for k, v in usd.items()}, strict=False)
    dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=vcfg["ch"], ch_mult=list(vcfg["ch_mult"]),
              num_res_blocks=vcfg["num_res_blocks"], attn_resolutions=[], dropout=0.0)
    model.first_stage_model = AutoencoderKL(ddconfig=dd, embed_dim=4)
    model.first_stage_model.load_state_dict(vsd, strict=True)
    model.cond_stage_model = FrozenCLIPImageEmbedder(**kcfg)
    model.cond_stage_model.load_state_dict(ksd, strict=True)
    model.proj_out = torch.nn.Linear(kcfg["width"], 768)
    model.proj_out.load_state_dict({"weight": proj_w, "bias": proj_b})
    model.learnable_vector.data.copy_(uc)
    model = model.to(dev).eval()
    with torch.no_grad():
        c = model.proj_out(model.get_learned_conditioning(d(ref_img)))
        post = model.encode_first_stage(d(image * mask))
        z_inp = model.get_first_stage_encoding(post.mode())
        samples, _ = PLMSSampler(model).sample(S=steps, conditioning=c, batch_size=B, shape=[4, 32, 32], verbose=False,
                                               unconditional_guidance_scale=scale,
                                               unconditional_conditioning=model.learnable_vector, eta=0.0, x_T=d(x_T),
                                               test_model_kwargs=dict(inpaint_image=z_inp, inpaint_mask=m_lat))
        img = model.decode_first_stage(samples)
        u8 = model.first_stage_model.decode_to_uint8(samples / model.scale_factor)
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    print(f"pipeline: cond rel-L2 {rel(c, c_ref):.3e}, latent z rel-L2 {rel(z_inp, z_ref):.3e}, "
          f"sampled latent PSNR {_psnr(samples, lat_ref):.1f} dB, image PSNR {_psnr(img, img_ref):.1f} dB")
    assert rel(c, c_ref) <= 1e-2 and rel(z_inp, z_ref) <= 1e-2
    assert _psnr(samples, lat_ref) >= 35.0 and _psnr(img, img_ref) >= 30.0
    assert u8.shape == (B, H, W, 3) and u8.dtype == torch.uint8


def test_test_bench_tool_with_png_io_overlapped(tmp_path):
    """tools/test_bench.py (BASELINE config C3 loop, scripts/inference_test_bench.py:295-397) on the CI-size networks with
    the real host I/O: PNG triples in the reference's folder layout are decoded by the loader threads, results are PNG-
    encoded by the writer threads; every request comes back as a readable 128x128 RGB file named <id:012d>.png."""
    import json
    import subprocess
    import sys
    import numpy as np
    from PIL import Image
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ds, out = tmp_path / "tb", tmp_path / "out"
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "test_bench.py"), "--small", "--requests", "10", "--batch", "4",
                        "--steps", "4", "--size", "128", "--dataset", str(ds), "--save", str(out), "--io-workers", "4"],
                       capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == 1 and line["requests_timed"] == 6 and line["images_per_sec"] > 0      # batch 0 (4 requests) builds the plans
    ids = np.load(ds / "id_list.npy").tolist()
    files = sorted(os.listdir(out / "results"))
    assert files == sorted(str(i).zfill(12) + ".png" for i in ids)
    img = np.asarray(Image.open(out / "results" / files[0]))
    assert img.shape == (128, 128, 3) and img.dtype == np.uint8 and img.std() > 0
