"""The drop-in switch (pbe_b200.install, pbe_b200/dropin.py): the reference's own `instantiate_from_config` on the UNEDITED
configs/v1.yaml and the unmodified scripts/inference.py resolve to this package's classes.  CPU tests: they stop where the
first device kernel would run (the product has no CPU path, and says so)."""
import importlib
import os
import sys
import types

import pytest
import torch

REF = os.environ.get("PBE_REFERENCE", "/root/reference")
HAVE_REF = os.path.isdir(os.path.join(REF, "ldm"))


@pytest.fixture
def installed():
    import pbe_b200
    added_path = False
    if HAVE_REF and REF not in sys.path:
        sys.path.insert(0, REF)
        added_path = True
    before = {k: sys.modules.get(k) for k in list(sys.modules) if k == "ldm" or k.startswith("ldm.")}
    pbe_b200.install()
    try:
        yield pbe_b200
    finally:
        pbe_b200.uninstall()
        for k in [k for k in sys.modules if (k == "ldm" or k.startswith("ldm.")) and k not in before]:
            if getattr(sys.modules[k], "__pbe_b200_stub__", False):
                sys.modules.pop(k, None)
        if added_path:
            sys.path.remove(REF)


SMALL_MODEL_CFG = {   # the `model:` node of configs/v1.yaml with the reference's target strings, narrowed for test speed
    "base_learning_rate": 1.0e-05,
    "target": "ldm.models.diffusion.ddpm.LatentDiffusion",
    "params": {
        "linear_start": 0.00085, "linear_end": 0.0120, "num_timesteps_cond": 1, "log_every_t": 200, "timesteps": 1000,
        "first_stage_key": "inpaint", "cond_stage_key": "image", "image_size": 64, "channels": 4,
        "cond_stage_trainable": True, "conditioning_key": "crossattn", "monitor": "val/loss_simple_ema",
        "u_cond_percent": 0.2, "scale_factor": 0.18215, "use_ema": False,
        "scheduler_config": {"target": "ldm.lr_scheduler.LambdaLinearScheduler", "params": {"warm_up_steps": [10000]}},
        "unet_config": {"target": "ldm.modules.diffusionmodules.openaimodel.UNetModel",
                        "params": {"image_size": 32, "in_channels": 9, "out_channels": 4, "model_channels": 64,
                                   "attention_resolutions": [4, 2, 1], "num_res_blocks": 1, "channel_mult": [1, 2, 4, 4],
                                   "num_heads": 8, "use_spatial_transformer": True, "transformer_depth": 1,
                                   "context_dim": 768, "use_checkpoint": True, "legacy": False,
                                   "add_conv_in_front_of_unet": False}},
        "first_stage_config": {"target": "ldm.models.autoencoder.AutoencoderKL",
                               "params": {"embed_dim": 4, "monitor": "val/rec_loss",
                                          "ddconfig": {"double_z": True, "z_channels": 4, "resolution": 256, "in_channels": 3,
                                                       "out_ch": 3, "ch": 32, "ch_mult": [1, 2, 4, 4], "num_res_blocks": 1,
                                                       "attn_resolutions": [], "dropout": 0.0},
                                          "lossconfig": {"target": "torch.nn.Identity"}}},
        "cond_stage_config": {"target": "ldm.modules.encoders.modules.FrozenCLIPImageEmbedder",
                              "params": {"width": 64, "layers": 2, "heads": 4, "mlp_dim": 128, "mapper_layers": 1}},
    },
}


def _check_model(model):
    from pbe_b200 import clip, diffusion, unet, vae
    assert type(model) is diffusion.LatentDiffusion
    assert type(model.model.diffusion_model) is unet.UNetModel
    assert type(model.first_stage_model) is vae.AutoencoderKL
    assert type(model.cond_stage_model) is clip.FrozenCLIPImageEmbedder
    assert model.learnable_vector.shape == (1, 1, 768) and model.proj_out.in_features == 1024
    keys = model.state_dict().keys()
    for k in ("model.diffusion_model.input_blocks.0.0.weight", "first_stage_model.decoder.conv_in.weight",
              "first_stage_model.quant_conv.weight", "cond_stage_model.final_ln.weight", "learnable_vector",
              "proj_out.weight", "betas"):
        assert k in keys, k


def test_install_standalone_serves_the_reference_module_paths(installed):
    """Without the reference on sys.path at all: the module paths the yaml and the scripts name resolve to this package."""
    from pbe_b200 import clip, diffusion, dropin, samplers, unet, vae
    assert dropin.installed()
    from ldm.models.diffusion.ddim import DDIMSampler            # scripts/inference.py:17
    from ldm.models.diffusion.plms import PLMSSampler            # scripts/inference.py:18
    assert PLMSSampler is samplers.PLMSSampler and DDIMSampler is samplers.DDIMSampler
    for path, cls in (("ldm.models.diffusion.ddpm.LatentDiffusion", diffusion.LatentDiffusion),
                      ("ldm.models.diffusion.latent_diffusion.LatentDiffusion", diffusion.LatentDiffusion),
                      ("ldm.modules.diffusionmodules.openaimodel.UNetModel", unet.UNetModel),
                      ("ldm.models.autoencoder.AutoencoderKL", vae.AutoencoderKL),
                      ("ldm.modules.encoders.modules.FrozenCLIPImageEmbedder", clip.FrozenCLIPImageEmbedder)):
        assert dropin.get_obj_from_str(path) is cls, path
    model = dropin.instantiate_from_config(SMALL_MODEL_CFG)
    _check_model(model)
    # a checkpoint-shaped state dict loads under the reference's key layout (scripts/inference.py:60-65)
    sd = {k: torch.full_like(v, 0.5) for k, v in model.state_dict().items()}
    sd["model_ema.decay"] = torch.zeros(())                         # training leftovers are reported, not fatal
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not missing and unexpected == ["model_ema.decay"]
    with model.ema_scope():
        pass
    with pytest.raises(KeyError, match="Expected key `target`"):
        dropin.instantiate_from_config({"params": {}})


def test_uninstall_restores_sys_modules():
    import pbe_b200
    from pbe_b200 import dropin
    marker = types.ModuleType("ldm.models.diffusion.plms")
    had = sys.modules.get("ldm.models.diffusion.plms")
    sys.modules["ldm.models.diffusion.plms"] = marker
    try:
        pbe_b200.install()
        assert sys.modules["ldm.models.diffusion.plms"] is importlib.import_module("pbe_b200.samplers")
        pbe_b200.uninstall()
        assert sys.modules["ldm.models.diffusion.plms"] is marker and not dropin.installed()
    finally:
        if had is None:
            sys.modules.pop("ldm.models.diffusion.plms", None)
        else:
            sys.modules["ldm.models.diffusion.plms"] = had
        for k in [k for k in sys.modules if k == "ldm" or k.startswith("ldm.")]:
            if getattr(sys.modules[k], "__pbe_b200_stub__", False):
                sys.modules.pop(k, None)


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not mounted (GPU box)")
def test_reference_instantiate_from_config_on_the_unedited_v1_yaml(installed):
    """The reference's OWN ldm/util.py:instantiate_from_config on the stock configs/v1.yaml (read with yaml: omegaconf is not
    in this image; the function only uses `in`, [] and .get) builds this package's classes at full v1 size."""
    import yaml
    util = importlib.import_module("ldm.util")
    assert util.__file__.startswith(REF)                          # the reference's resolver, not ours
    cfg = yaml.safe_load(open(os.path.join(REF, "configs", "v1.yaml")))
    assert cfg["model"]["target"] == "ldm.models.diffusion.ddpm.LatentDiffusion"
    model = util.instantiate_from_config(cfg["model"])
    _check_model(model)
    unet = model.model.diffusion_model
    assert sum(p.numel() for p in unet.parameters()) == 859_535_364 - 0   # the v1.yaml U-Net
    assert model.first_stage_model.ch_mult == (1, 2, 4, 4) and model.cond_stage_model.layers == 24
    assert model.scale_factor == 0.18215 and model.num_timesteps == 1000


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not mounted (GPU box)")
def test_unmodified_inference_script_runs_into_this_package(installed, tmp_path, monkeypatch):
    """scripts/inference.py, byte for byte, with test.sh's arguments: argument parsing, OmegaConf.load, load_model_from_config
    (instantiate + load_state_dict), sampler construction and the PIL pre-processing all run; the first model call
    (get_learned_conditioning -> the CLIP front-end) is this package's CUDA path, which refuses to run on this CPU-only
    box -- there is no fallback to fall into.  Third-party packages the image lacks (omegaconf, imwatermark, diffusers,
    pytorch_lightning, clip) and the safety checker's network download are stubbed; nothing of `ldm` is."""
    import yaml

    class Node(dict):
        __getattr__ = dict.__getitem__

    def to_node(o):
        if isinstance(o, dict):
            return Node({k: to_node(v) for k, v in o.items()})
        if isinstance(o, list):
            return [to_node(v) for v in o]
        return o

    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        monkeypatch.setitem(sys.modules, name, m)
        return m

    mod("omegaconf", OmegaConf=type("OmegaConf", (), {"load": staticmethod(lambda p: to_node(yaml.safe_load(open(p))))}))

    class WatermarkEncoder:
        def set_watermark(self, *a, **k):
            pass

    mod("imwatermark", WatermarkEncoder=WatermarkEncoder)
    mod("pytorch_lightning", seed_everything=lambda s: torch.manual_seed(s))
    mod("clip")
    fake = type("Pretrained", (), {"from_pretrained": classmethod(lambda cls, *a, **k: cls())})
    mod("diffusers")
    mod("diffusers.pipelines")
    mod("diffusers.pipelines.stable_diffusion")
    mod("diffusers.pipelines.stable_diffusion.safety_checker", StableDiffusionSafetyChecker=fake)
    import transformers
    monkeypatch.setattr(transformers, "AutoFeatureExtractor", fake, raising=False)
    # the checkpoint: an empty state dict (load_state_dict(strict=False) then reports every key missing and goes on)
    monkeypatch.setattr(torch, "load", lambda *a, **k: {"state_dict": {}})
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, device=None: self)       # no GPU in this container
    monkeypatch.chdir(tmp_path)
    argv = ["inference.py", "--plms", "--outdir", str(tmp_path / "results"), "--config", os.path.join(REF, "configs/v1.yaml"),
            "--ckpt", "checkpoints/model.ckpt", "--image_path", os.path.join(REF, "examples/image/example_1.png"),
            "--mask_path", os.path.join(REF, "examples/mask/example_1.png"),
            "--reference_path", os.path.join(REF, "examples/reference/example_1.jpg"), "--seed", "321", "--scale", "5"]
    monkeypatch.setattr(sys, "argv", argv)
    spec = importlib.util.spec_from_file_location("pbe_reference_inference_script", os.path.join(REF, "scripts/inference.py"))
    script = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(script)
    from pbe_b200 import samplers
    assert script.PLMSSampler is samplers.PLMSSampler and script.DDIMSampler is samplers.DDIMSampler
    with pytest.raises(RuntimeError, match="runs only on a CUDA"):
        script.main()
    assert (tmp_path / "results" / "results").is_dir()     # the script got past model + sampler construction
