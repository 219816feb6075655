"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the authoring container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py small      # seconds..a minute
    python tests/golden/make_golden.py v1         # one full-size U-Net call (~10 s) + PLMS-50 C1 trajectory (~10 min)
    python tests/golden/make_golden.py v1_ddim20  # DDIM-20 C4 trajectory at v1.yaml size (~4 min)
    python tests/golden/make_golden.py vae        # VAE decode: small config + the v1.yaml decoder on a 16x16 latent
    python tests/golden/make_golden.py clip       # conditioning front-end: live transformers CLIPVisionModel + reference mapper

Inputs and weights are regenerated from seeds (oracle.unet_ref.make_state_dict, oracle.sampler_ref.synthetic_request),
so only outputs are stored.  Every file is written as float32 .npy; golden_index.json records shapes, seeds and sha256.
"""
import hashlib, json, os, sys, time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import unet_ref as U, sampler_ref as S, reference_bridge as R, vae_ref as V, clip_ref as K

OUT = os.path.dirname(os.path.abspath(__file__))
INDEX = os.path.join(OUT, "golden_index.json")


def save(name, t, meta, index):
    a = t.detach().cpu().numpy().astype(np.float32)
    np.save(os.path.join(OUT, name + ".npy"), a)
    index[name] = dict(shape=list(a.shape), sha256=hashlib.sha256(a.tobytes()).hexdigest(), **meta)
    print("wrote", name, a.shape, flush=True)


def run_sampler(kind, model, req, Sn, hw, B, scale=5.0):
    smp = R.reference_sampler(kind, model)
    kw = dict(S=Sn, conditioning=req["c"], batch_size=B, shape=[4, hw, hw], verbose=False,
              unconditional_guidance_scale=scale, unconditional_conditioning=req["uc"].expand(B, 1, 768), eta=0.0,
              x_T=req["x_T"], test_model_kwargs=dict(images_inpaint=req["z_inpaint"], images_mask=req["mask"]))
    if kind == "ddim":
        kw["disable_tqdm"] = True
    out, inter = smp.sample(**kw)
    return out, inter


def main(which):
    index = json.load(open(INDEX)) if os.path.exists(INDEX) else {}
    torch.set_num_threads(os.cpu_count())
    if which == "small":
        cfg = U.SMALL_CFG
        sd = U.make_state_dict(cfg, 321)
        ref = R.build_reference_unet(cfg, sd)
        index["state_dict_keys"] = dict(keys=sorted(ref.state_dict().keys()), n=len(ref.state_dict()))
        B, hw = 2, 32
        req = S.synthetic_request(B, hw, hw, seed=321)
        x9 = torch.cat((req["x_T"], req["z_inpaint"], req["mask"]), 1)
        x_in = torch.cat([x9] * 2)
        c_in = torch.cat((req["uc"].expand(B, 1, 768), req["c"]))
        for tval in (981, 1):
            t = torch.full((2 * B,), tval, dtype=torch.int64)
            with torch.no_grad():
                e = ref(x_in, t, context=c_in)
            save(f"small_unet_eps_t{tval}", e, dict(cfg="SMALL_CFG", weight_seed=321, request_seed=321, B=B, hw=hw, t=tval,
                                                  source="reference UNetModel.forward fp32 CPU"), index)
        model = R.StubLatentDiffusion(ref)
        out, _ = run_sampler("plms", model, req, 8, hw, B)
        save("small_plms8_final", out, dict(cfg="SMALL_CFG", S=8, scale=5.0, source="reference PLMSSampler.sample"), index)
        out, _ = run_sampler("ddim", model, req, 5, hw, B)
        save("small_ddim5_final", out, dict(cfg="SMALL_CFG", S=5, scale=5.0, source="reference DDIMSampler.sample"), index)
        out, _ = run_sampler("plms", model, req, 50, hw, B)
        save("small_plms50_final", out, dict(cfg="SMALL_CFG", S=50, scale=5.0, source="reference PLMSSampler.sample"), index)
    elif which == "v1":
        cfg = U.V1_CFG
        t0 = time.time()
        sd = U.make_state_dict(cfg, 321)
        ref = R.build_reference_unet(cfg, sd)
        print("built v1 reference", time.time() - t0, flush=True)
        B, hw = 1, 64
        req = S.synthetic_request(B, hw, hw, seed=321)
        x9 = torch.cat((req["x_T"], req["z_inpaint"], req["mask"]), 1)
        x_in = torch.cat([x9] * 2)
        c_in = torch.cat((req["uc"].expand(B, 1, 768), req["c"]))
        for tval in (981, 1):
            t = torch.full((2 * B,), tval, dtype=torch.int64)
            t0 = time.time()
            with torch.no_grad():
                e = ref(x_in, t, context=c_in)
            print("v1 call", time.time() - t0, "s", flush=True)
            save(f"v1_unet_eps_t{tval}", e, dict(cfg="V1_CFG", weight_seed=321, request_seed=321, B=B, hw=hw, t=tval,
                                               source="reference UNetModel.forward fp32 CPU"), index)
        json.dump(index, open(INDEX, "w"), indent=1, sort_keys=True)
        model = R.StubLatentDiffusion(ref)
        t0 = time.time()
        out, inter = run_sampler("plms", model, req, 50, hw, B)
        print("v1 plms50", time.time() - t0, "s, unet calls", model.calls, flush=True)
        save("v1_plms50_final", out, dict(cfg="V1_CFG", S=50, scale=5.0, B=1, hw=64, unet_calls=model.calls,
                                          cpu_seconds=time.time() - t0, cpu_threads=torch.get_num_threads(),
                                          source="reference PLMSSampler.sample (BASELINE config C1)"), index)
    elif which == "v1_ddim20":
        # BASELINE config C4: DDIM 20 steps, batch 1, 64x64 latent, scale 5 (test.sh:1-9; ddim.py:136-242) -- 20 CFG U-Net calls
        cfg = U.V1_CFG
        sd = U.make_state_dict(cfg, 321)
        ref = R.build_reference_unet(cfg, sd)
        B, hw = 1, 64
        req = S.synthetic_request(B, hw, hw, seed=321)
        model = R.StubLatentDiffusion(ref)
        t0 = time.time()
        out, inter = run_sampler("ddim", model, req, 20, hw, B)
        print("v1 ddim20", time.time() - t0, "s, unet calls", model.calls, flush=True)
        save("v1_ddim20_final", out, dict(cfg="V1_CFG", S=20, scale=5.0, B=1, hw=64, unet_calls=model.calls,
                                          cpu_seconds=time.time() - t0, cpu_threads=torch.get_num_threads(),
                                          source="reference DDIMSampler.sample (BASELINE config C4)"), index)
    elif which == "vae":
        for tag, cfg, B, hw in (("small", V.SMALL_VAE_CFG, 2, 16), ("v1", V.V1_VAE_CFG, 1, 16)):
            sd = V.make_state_dict(cfg, 321)
            dec = R.build_reference_vae_decode(cfg, sd)
            enc = R.build_reference_vae_encode(cfg, sd)
            if tag == "v1":
                keys = sorted(dec.state_dict_keys + enc.state_dict_keys)
                index["vae_state_dict_keys"] = dict(keys=keys, n=len(keys))
            f = 2 ** (len(cfg["ch_mult"]) - 1)
            x = V.synthetic_images(B, 16 * f, 16 * f, seed=321)
            mom = enc(x)
            save(f"{tag}_vae_encode_moments", mom, dict(cfg=("SMALL_VAE_CFG" if tag == "small" else "V1_VAE_CFG"), weight_seed=321,
                                                        image_seed=321, B=B, hw=16 * f,
                                                        source="reference quant_conv(Encoder(x)) fp32 CPU"), index)
            z = V.synthetic_latents(B, hw, hw, seed=321)
            t0 = time.time()
            img = dec(z)
            print(tag, "vae decode", time.time() - t0, "s", flush=True)
            save(f"{tag}_vae_decode", img, dict(cfg=("SMALL_VAE_CFG" if tag == "small" else "V1_VAE_CFG"), weight_seed=321,
                                                latent_seed=321, B=B, hw=hw,
                                                source="reference Decoder(post_quant_conv(z)) fp32 CPU"), index)
    elif which == "clip":
        for tag, cfg, B in (("small", K.SMALL_CLIP_CFG, 2), ("v1", K.V1_CLIP_CFG, 1)):
            sd = K.make_state_dict(cfg, 321)
            enc = R.build_reference_clip_embedder(cfg, sd)
            if tag == "v1":
                index["clip_state_dict_keys"] = dict(keys=enc.state_dict_keys, n=len(enc.state_dict_keys))
            x = K.synthetic_exemplars(B, cfg["image_size"], seed=321)
            t0 = time.time()
            z = enc(x)
            print(tag, "clip encode", time.time() - t0, "s", flush=True)
            import transformers
            save(f"{tag}_clip_embed", z, dict(cfg=("SMALL_CLIP_CFG" if tag == "small" else "V1_CLIP_CFG"), weight_seed=321,
                                              image_seed=321, B=B, transformers=transformers.__version__,
                                              source="transformers CLIPVisionModel.pooler_output -> reference xf mapper -> "
                                                     "final_ln, fp32 CPU"), index)
    json.dump(index, open(INDEX, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "small")
