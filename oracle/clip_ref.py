"""ORACLE — test infrastructure only (never imported by the product path ``pbe_b200/``).

Plain-PyTorch fp32 restatement of ``FrozenCLIPImageEmbedder.forward`` (reference ldm/modules/encoders/modules.py:160-166):
the CLIP vision tower lives in a third-party dependency that is not vendored in /root/reference — ``transformers``
(pinned ``transformers==4.19.2``, environment.yaml:27; 5.5.0 in this image), class ``CLIPVisionModel`` /
``CLIPVisionTransformer`` (models/clip/modeling_clip.py) — so its published algorithm is restated here (ViT with a
class token, learned positions, pre-LN blocks, quick_gelu MLP, post-LN of the class token) and pinned against the live
``transformers`` implementation; the mapper and final LayerNorm are the reference's own ldm/modules/encoders/xf.py.
Pinned by ``tests/test_oracle_pinned.py`` (live import) and the goldens of ``tests/golden/make_golden.py clip``.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

V1_CLIP_CFG = dict(  # openai/clip-vit-large-patch14 vision tower + Transformer(1, 1024, 5, 1), modules.py:141-151
    image_size=224, patch_size=14, width=1024, layers=24, heads=16, mlp_dim=4096, mapper_layers=5)

SMALL_CLIP_CFG = dict(  # same topology, tiny: CPU-fast parity config (5 x 5 patches -> 26 tokens)
    image_size=70, patch_size=14, width=128, layers=2, heads=2, mlp_dim=512, mapper_layers=2)

VM = "transformer.vision_model."


def param_shapes(cfg) -> Dict[str, tuple]:
    C, P, F_ = cfg["width"], (cfg["image_size"] // cfg["patch_size"]) ** 2, cfg["mlp_dim"]
    s: Dict[str, tuple] = {}

    def lin(p, o, i):
        s[p + ".weight"] = (o, i)
        s[p + ".bias"] = (o,)

    def norm(p):
        s[p + ".weight"] = (C,)
        s[p + ".bias"] = (C,)

    s[VM + "embeddings.class_embedding"] = (C,)
    s[VM + "embeddings.patch_embedding.weight"] = (C, 3, cfg["patch_size"], cfg["patch_size"])
    s[VM + "embeddings.position_embedding.weight"] = (P + 1, C)
    norm(VM + "pre_layrnorm")
    for i in range(cfg["layers"]):
        lp = f"{VM}encoder.layers.{i}."
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            lin(lp + "self_attn." + n, C, C)
        norm(lp + "layer_norm1")
        lin(lp + "mlp.fc1", F_, C)
        lin(lp + "mlp.fc2", C, F_)
        norm(lp + "layer_norm2")
    norm(VM + "post_layernorm")
    norm("final_ln")
    for j in range(cfg["mapper_layers"]):
        mp = f"mapper.resblocks.{j}."
        lin(mp + "attn.c_qkv", 3 * C, C)
        lin(mp + "attn.c_proj", C, C)
        norm(mp + "ln_1")
        lin(mp + "mlp.c_fc", 4 * C, C)
        lin(mp + "mlp.c_proj", C, 4 * C)
        norm(mp + "ln_2")
    return s


def make_state_dict(cfg, seed: int = 321) -> Dict[str, torch.Tensor]:
    """Deterministic, name-keyed random weights (same scheme as oracle/unet_ref.py::make_state_dict)."""
    import zlib
    sd = {}
    for name, shape in param_shapes(cfg).items():
        g = torch.Generator(device="cpu").manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
        leaf = name.rsplit(".", 1)[1]
        if name.endswith("class_embedding") or name.endswith("position_embedding.weight"):
            t = 0.3 * torch.randn(shape, generator=g)
        elif len(shape) == 1 and leaf == "weight":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        sd[name] = t
    return sd


def _ln(sd, p, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def vision_pooled(sd, cfg, image: torch.Tensor) -> torch.Tensor:
    """transformers CLIPVisionTransformer.forward -> pooler_output (modeling_clip.py: CLIPVisionEmbeddings, pre_layrnorm,
    CLIPEncoderLayer x N with CLIPAttention (scale = head_dim^-0.5 applied to q) and CLIPMLP (quick_gelu),
    post_layernorm(last_hidden_state[:, 0, :]))."""
    C, H = cfg["width"], cfg["heads"]
    d = C // H
    B = image.shape[0]
    patch = F.conv2d(image, sd[VM + "embeddings.patch_embedding.weight"], stride=cfg["patch_size"])
    patch = patch.flatten(2).transpose(1, 2)
    cls = sd[VM + "embeddings.class_embedding"].expand(B, 1, C)
    x = torch.cat([cls, patch], dim=1) + sd[VM + "embeddings.position_embedding.weight"]
    x = _ln(sd, VM + "pre_layrnorm", x)
    N = x.shape[1]
    for i in range(cfg["layers"]):
        lp = f"{VM}encoder.layers.{i}."
        h = _ln(sd, lp + "layer_norm1", x)
        q = _lin(sd, lp + "self_attn.q_proj", h) * d ** -0.5
        k = _lin(sd, lp + "self_attn.k_proj", h)
        v = _lin(sd, lp + "self_attn.v_proj", h)
        split = lambda t: t.view(B, N, H, d).transpose(1, 2)
        w = torch.softmax(split(q) @ split(k).transpose(-1, -2), dim=-1)
        a = (w @ split(v)).transpose(1, 2).reshape(B, N, C)
        x = x + _lin(sd, lp + "self_attn.out_proj", a)
        h = _ln(sd, lp + "layer_norm2", x)
        h = _lin(sd, lp + "mlp.fc1", h)
        h = h * torch.sigmoid(1.702 * h)                       # quick_gelu
        x = x + _lin(sd, lp + "mlp.fc2", h)
    return _ln(sd, VM + "post_layernorm", x[:, 0, :])


def mapper(sd, cfg, z: torch.Tensor) -> torch.Tensor:
    """xf.py Transformer(n_ctx=1, width, layers, heads=1).forward on z [B, n_ctx, C] (xf.py:43-66,86-104,125-130)."""
    for j in range(cfg["mapper_layers"]):
        mp = f"mapper.resblocks.{j}."
        h = _ln(sd, mp + "ln_1", z)
        qkv = _lin(sd, mp + "attn.c_qkv", h)
        bs, n_ctx, width = qkv.shape
        heads = 1
        attn_ch = width // heads // 3
        scale = 1 / math.sqrt(math.sqrt(attn_ch))
        qkv = qkv.view(bs, n_ctx, heads, -1)
        q, k, v = torch.split(qkv, attn_ch, dim=-1)
        weight = torch.softmax(torch.einsum("bthc,bshc->bhts", q * scale, k * scale).float(), dim=-1)
        a = torch.einsum("bhts,bshc->bthc", weight, v).reshape(bs, n_ctx, -1)
        z = z + _lin(sd, mp + "attn.c_proj", a)
        h = _ln(sd, mp + "ln_2", z)
        z = z + _lin(sd, mp + "mlp.c_proj", F.gelu(_lin(sd, mp + "mlp.c_fc", h)))
    return z


def encode(sd, cfg, image: torch.Tensor) -> torch.Tensor:
    """FrozenCLIPImageEmbedder.forward, modules.py:160-166 -> [B, 1, width]."""
    z = vision_pooled(sd, cfg, image).unsqueeze(1)
    return _ln(sd, "final_ln", mapper(sd, cfg, z))


def synthetic_exemplars(B: int, size: int, seed: int = 321) -> torch.Tensor:
    """CLIP-normalised exemplar crops: roughly unit-variance, spatially smooth + noise."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    low = F.interpolate(torch.randn(B, 3, max(size // 14, 1), max(size // 14, 1), generator=g), size=(size, size),
                        mode="bilinear", align_corners=False)
    return low + 0.3 * torch.randn(B, 3, size, size, generator=g)
