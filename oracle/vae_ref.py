"""ORACLE — test infrastructure only (never imported by the product path ``pbe_b200/``).

Plain-PyTorch fp32 restatement of ``AutoencoderKL.decode`` (reference ldm/models/autoencoder.py:66-69) =
``Decoder.forward`` (ldm/modules/diffusionmodules/model.py:542-580) after ``post_quant_conv``, written functionally over
a state dict with the reference's own key names.  Pinned against the real reference ``Decoder`` module by
``tests/test_oracle_pinned.py`` (live import, authoring container only) and against the committed golden vectors in
``tests/golden/`` (``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

V1_VAE_CFG = dict(  # configs/v1.yaml:48-67
    embed_dim=4, z_channels=4, ch=128, out_ch=3, in_channels=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2)

SMALL_VAE_CFG = dict(  # same topology (three levels), narrower: CPU-fast parity config
    embed_dim=4, z_channels=4, ch=64, out_ch=3, in_channels=3, ch_mult=(1, 2, 2), num_res_blocks=1)


def param_shapes(cfg) -> Dict[str, tuple]:
    """post_quant_conv (autoencoder.py:37) + Decoder.__init__ (model.py:475-540) for attn_resolutions=[]."""
    s: Dict[str, tuple] = {}

    def conv(p, o, i, k):
        s[p + ".weight"] = (o, i, k, k)
        s[p + ".bias"] = (o,)

    def norm(p, c):
        s[p + ".weight"] = (c,)
        s[p + ".bias"] = (c,)

    def res(p, cin, cout):
        norm(p + ".norm1", cin); conv(p + ".conv1", cout, cin, 3)
        norm(p + ".norm2", cout); conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".nin_shortcut", cout, cin, 1)

    ch, mult, L = cfg["ch"], cfg["ch_mult"], len(cfg["ch_mult"])
    conv("post_quant_conv", cfg["z_channels"], cfg["embed_dim"], 1)
    block_in = ch * mult[L - 1]
    conv("decoder.conv_in", block_in, cfg["z_channels"], 3)
    res("decoder.mid.block_1", block_in, block_in)
    norm("decoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"decoder.mid.attn_1.{n}", block_in, block_in, 1)
    res("decoder.mid.block_2", block_in, block_in)
    for lvl in reversed(range(L)):
        block_out = ch * mult[lvl]
        for i in range(cfg["num_res_blocks"] + 1):
            res(f"decoder.up.{lvl}.block.{i}", block_in, block_out)
            block_in = block_out
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", block_in, block_in, 3)
    norm("decoder.norm_out", block_in)
    conv("decoder.conv_out", cfg["out_ch"], block_in, 3)
    # quant_conv (autoencoder.py:36) + Encoder.__init__ (model.py:370-438), double_z=True
    conv("encoder.conv_in", ch, cfg["in_channels"], 3)
    block_in = ch
    for lvl in range(L):
        block_out = ch * mult[lvl]
        for i in range(cfg["num_res_blocks"]):
            res(f"encoder.down.{lvl}.block.{i}", block_in, block_out)
            block_in = block_out
        if lvl != L - 1:
            conv(f"encoder.down.{lvl}.downsample.conv", block_in, block_in, 3)
    res("encoder.mid.block_1", block_in, block_in)
    norm("encoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"encoder.mid.attn_1.{n}", block_in, block_in, 1)
    res("encoder.mid.block_2", block_in, block_in)
    norm("encoder.norm_out", block_in)
    conv("encoder.conv_out", 2 * cfg["z_channels"], block_in, 3)
    conv("quant_conv", 2 * cfg["embed_dim"], 2 * cfg["z_channels"], 1)
    return s


def make_state_dict(cfg, seed: int = 321) -> Dict[str, torch.Tensor]:
    """Deterministic, name-keyed random weights (same scheme as oracle/unet_ref.py::make_state_dict)."""
    import zlib
    sd = {}
    for name, shape in param_shapes(cfg).items():
        g = torch.Generator(device="cpu").manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
        leaf = name.rsplit(".", 1)[1]
        if len(shape) == 1 and leaf == "weight":
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


def _swish(x):
    """model.py:35-37."""
    return x * torch.sigmoid(x)


def _norm(sd, p, x):
    """Normalize = GroupNorm(32, eps=1e-6), model.py:40-41."""
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)


def _conv(sd, p, x, pad):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=pad)


def resnet_block(sd, p, x):
    """ResnetBlock.forward with temb=None, dropout 0 (model.py:123-143)."""
    h = _conv(sd, p + ".conv1", _swish(_norm(sd, p + ".norm1", x)), 1)
    h = _conv(sd, p + ".conv2", _swish(_norm(sd, p + ".norm2", h)), 1)
    if p + ".nin_shortcut.weight" in sd:
        x = _conv(sd, p + ".nin_shortcut", x, 0)
    return x + h


def attn_block(sd, p, x):
    """AttnBlock.forward (model.py:166-182): single head over all h*w positions, scale c^-1/2."""
    h_ = _norm(sd, p + ".norm", x)
    q, k, v = (_conv(sd, f"{p}.{n}", h_, 0) for n in ("q", "k", "v"))
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, hh * ww)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(sd, p + ".proj_out", h_, 0)


def decode(sd: Dict[str, torch.Tensor], cfg, z: torch.Tensor, taps: dict | None = None) -> torch.Tensor:
    """AutoencoderKL.decode (autoencoder.py:66-69) -> Decoder.forward (model.py:542-580). z: [B, embed_dim, h, w]."""
    L = len(cfg["ch_mult"])
    h = _conv(sd, "post_quant_conv", z, 0)
    h = _conv(sd, "decoder.conv_in", h, 1)
    h = resnet_block(sd, "decoder.mid.block_1", h)
    h = attn_block(sd, "decoder.mid.attn_1", h)
    h = resnet_block(sd, "decoder.mid.block_2", h)
    if taps is not None:
        taps["mid"] = h
    for lvl in reversed(range(L)):
        for i in range(cfg["num_res_blocks"] + 1):
            h = resnet_block(sd, f"decoder.up.{lvl}.block.{i}", h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")     # Upsample.forward, model.py:55-59
            h = _conv(sd, f"decoder.up.{lvl}.upsample.conv", h, 1)
        if taps is not None:
            taps[f"up{lvl}"] = h
    h = _swish(_norm(sd, "decoder.norm_out", h))
    return _conv(sd, "decoder.conv_out", h, 1)


def encode_moments(sd: Dict[str, torch.Tensor], cfg, x: torch.Tensor) -> torch.Tensor:
    """AutoencoderKL.encode up to the posterior parameters (autoencoder.py:56-61): quant_conv(Encoder.forward(x)),
    model.py:440-471.  x: [B, in_channels, H, W] -> moments [B, 2*embed_dim, H/f, W/f] (mean | logvar)."""
    L = len(cfg["ch_mult"])
    h = _conv(sd, "encoder.conv_in", x, 1)
    for lvl in range(L):
        for i in range(cfg["num_res_blocks"]):
            h = resnet_block(sd, f"encoder.down.{lvl}.block.{i}", h)
        if lvl != L - 1:
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)      # Downsample.forward, model.py:73-77
            h = F.conv2d(h, sd[f"encoder.down.{lvl}.downsample.conv.weight"], sd[f"encoder.down.{lvl}.downsample.conv.bias"],
                         stride=2, padding=0)
    h = resnet_block(sd, "encoder.mid.block_1", h)
    h = attn_block(sd, "encoder.mid.attn_1", h)
    h = resnet_block(sd, "encoder.mid.block_2", h)
    h = _swish(_norm(sd, "encoder.norm_out", h))
    h = _conv(sd, "encoder.conv_out", h, 1)
    return _conv(sd, "quant_conv", h, 0)


def synthetic_images(B: int, H: int, W: int, seed: int = 321) -> torch.Tensor:
    """Images in [-1, 1] as encode_first_stage sees them (smooth + noise, so that the statistics are image-like)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    low = F.interpolate(torch.randn(B, 3, max(H // 16, 1), max(W // 16, 1), generator=g), size=(H, W), mode="bilinear",
                        align_corners=False)
    return (0.6 * low + 0.15 * torch.randn(B, 3, H, W, generator=g)).clamp(-1, 1)


def synthetic_latents(B: int, h: int, w: int, seed: int = 321) -> torch.Tensor:
    """Latents as decode_first_stage sees them: sampled z / scale_factor, i.e. roughly N(0, 1/0.18215^2 * 0.8^2)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(B, 4, h, w, generator=g) * 4.5
