"""Drop-in for ``ldm.models.autoencoder.AutoencoderKL`` (reference autoencoder.py:14-69) — the first "next" row of
SURVEY.md §8f: ``LatentDiffusion.encode_first_stage`` (latent_diffusion.py:571-610) encodes the masked image once per
request before the PLMS/DDIM loop and ``decode_first_stage`` (latent_diffusion.py:444-508) turns the sampled latents into
images right after it.

Same constructor keywords as the reference (``ddconfig``, ``embed_dim``; ``lossconfig`` & co. are accepted and ignored),
same state-dict keys (``encoder.*``, ``quant_conv.*``, ``decoder.*``, ``post_quant_conv.*`` — a Paint-by-Example
checkpoint's ``first_stage_model.*`` entries load with ``strict=False``, ``loss.*`` is not used), same ``encode(x)`` /
``decode(z)`` contracts.  Each is one call into the sm_100a library (``pbe_vae_encode`` / ``pbe_vae_decode`` in
include/pbe_b200.h); there is no PyTorch/CPU fallback.
"""
from __future__ import annotations

import ctypes
from typing import Dict

import torch
import torch.nn as nn

from . import _lib
from .unet import _Node


class PbeVaeConfig(ctypes.Structure):
    _fields_ = [("embed_dim", ctypes.c_int32), ("z_channels", ctypes.c_int32), ("ch", ctypes.c_int32),
                ("out_ch", ctypes.c_int32), ("num_levels", ctypes.c_int32), ("ch_mult", ctypes.c_int32 * 8),
                ("num_res_blocks", ctypes.c_int32), ("in_channels", ctypes.c_int32)]


def vae_decoder_param_shapes(embed_dim, z_channels, ch, out_ch, ch_mult, num_res_blocks) -> Dict[str, tuple]:
    """State-dict keys and shapes of ``post_quant_conv`` + ``Decoder`` for attn_resolutions=[] (mirrors
    Decoder.__init__, ldm/modules/diffusionmodules/model.py:475-540, and autoencoder.py:37)."""
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

    conv("post_quant_conv", z_channels, embed_dim, 1)
    L = len(ch_mult)
    block_in = ch * ch_mult[L - 1]
    conv("decoder.conv_in", block_in, z_channels, 3)
    res("decoder.mid.block_1", block_in, block_in)
    norm("decoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"decoder.mid.attn_1.{n}", block_in, block_in, 1)
    res("decoder.mid.block_2", block_in, block_in)
    for lvl in reversed(range(L)):
        block_out = ch * ch_mult[lvl]
        for i in range(num_res_blocks + 1):
            res(f"decoder.up.{lvl}.block.{i}", block_in, block_out)
            block_in = block_out
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", block_in, block_in, 3)
    norm("decoder.norm_out", block_in)
    conv("decoder.conv_out", out_ch, block_in, 3)
    return s


def vae_encoder_param_shapes(embed_dim, z_channels, ch, in_channels, ch_mult, num_res_blocks) -> Dict[str, tuple]:
    """State-dict keys and shapes of ``Encoder`` (double_z=True, attn_resolutions=[]) + ``quant_conv`` (mirrors
    Encoder.__init__, ldm/modules/diffusionmodules/model.py:370-438, and autoencoder.py:36)."""
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

    L = len(ch_mult)
    conv("encoder.conv_in", ch, in_channels, 3)
    block_in = ch
    for lvl in range(L):
        block_out = ch * ch_mult[lvl]
        for i in range(num_res_blocks):
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
    conv("encoder.conv_out", 2 * z_channels, block_in, 3)
    conv("quant_conv", 2 * embed_dim, 2 * z_channels, 1)
    return s


class DiagonalGaussianDistribution:
    """Host-side mirror of ldm/modules/distributions/distributions.py:24-60 over the moments ``pbe_vae_encode`` returns
    (a handful of elementwise ops on a [B, 2*embed, h, w] tensor; ``sample`` draws its noise on the CPU generator and
    moves it to the device exactly as the reference does, distributions.py:35-37)."""

    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if self.deterministic:
            self.var = self.std = torch.zeros_like(self.mean)

    def sample(self):
        return self.mean + self.std * torch.randn(self.mean.shape).to(device=self.parameters.device)

    def mode(self):
        return self.mean

    def kl(self, other=None):
        if self.deterministic:
            return torch.Tensor([0.])
        if other is None:
            return 0.5 * torch.sum(torch.pow(self.mean, 2) + self.var - 1.0 - self.logvar, dim=[1, 2, 3])
        return 0.5 * torch.sum(torch.pow(self.mean - other.mean, 2) / other.var + self.var / other.var - 1.0
                               - self.logvar + other.logvar, dim=[1, 2, 3])


class AutoencoderKL(nn.Module):
    def __init__(self, ddconfig, embed_dim, lossconfig=None, ckpt_path=None, ignore_keys=(), image_key="image",
                 colorize_nlabels=None, monitor=None, **ignored):
        super().__init__()
        dd = dict(ddconfig)
        if list(dd.get("attn_resolutions", [])):
            raise NotImplementedError("pbe_b200.AutoencoderKL supports attn_resolutions=[] only (configs/v1.yaml:66)")
        if dd.get("dropout", 0.0):
            raise NotImplementedError("inference only: dropout must be 0")
        for k, want in (("tanh_out", False), ("give_pre_end", False), ("use_linear_attn", False), ("resamp_with_conv", True)):
            if dd.get(k, want) != want:
                raise NotImplementedError(f"pbe_b200.AutoencoderKL supports {k}={want!r} only")
        if dd.get("attn_type", "vanilla") != "vanilla":
            raise NotImplementedError("pbe_b200.AutoencoderKL supports attn_type='vanilla' only")
        if ckpt_path is not None:
            raise NotImplementedError("load weights with load_state_dict (ckpt_path is not supported)")
        if not dd.get("double_z", True):
            raise NotImplementedError("pbe_b200.AutoencoderKL supports double_z=True only (configs/v1.yaml:54)")
        self.embed_dim = int(embed_dim)
        self.in_channels = int(dd.get("in_channels", 3))
        self.z_channels = int(dd["z_channels"])
        self.ch = int(dd["ch"])
        self.out_ch = int(dd["out_ch"])
        self.ch_mult = tuple(int(m) for m in dd["ch_mult"])
        self.num_res_blocks = int(dd["num_res_blocks"])
        self._shapes = vae_encoder_param_shapes(self.embed_dim, self.z_channels, self.ch, self.in_channels, self.ch_mult,
                                                self.num_res_blocks)
        self._shapes.update(vae_decoder_param_shapes(self.embed_dim, self.z_channels, self.ch, self.out_ch, self.ch_mult,
                                                     self.num_res_blocks))
        for name, shape in self._shapes.items():
            parts = name.split(".")
            node = self
            for part in parts[:-1]:
                if not hasattr(node, part):
                    node.add_module(part, _Node())
                node = getattr(node, part)
            node.register_parameter(parts[-1], nn.Parameter(torch.zeros(shape), requires_grad=False))
        self._weights_epoch = 0
        self._engine = None
        self._engine_version = None
        self._engine_device = None

    # ------------------------------------------------------------------------------------------------------------
    def mark_weights_changed(self) -> None:
        self._weights_epoch += 1

    def _load_from_state_dict(self, *args, **kwargs):
        self._weights_epoch += 1
        return super()._load_from_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._weights_epoch += 1
        return super()._apply(fn, *args, **kwargs)

    def _destroy_engine(self):
        if self._engine is not None:
            _lib.load().pbe_vae_destroy(self._engine)
            self._engine = None

    def __del__(self):
        try:
            self._destroy_engine()
        except Exception:
            pass

    def _ensure_engine(self, device: torch.device):
        if device.type != "cuda":
            raise RuntimeError("pbe_b200.AutoencoderKL runs only on a CUDA (sm_100a) device: no CPU fallback exists")
        if self._engine is not None and self._engine_version == self._weights_epoch and self._engine_device == device:
            return
        self._destroy_engine()
        lib = _lib.load()
        cfg = PbeVaeConfig()
        cfg.embed_dim, cfg.z_channels, cfg.ch, cfg.out_ch = self.embed_dim, self.z_channels, self.ch, self.out_ch
        cfg.num_levels, cfg.num_res_blocks, cfg.in_channels = len(self.ch_mult), self.num_res_blocks, self.in_channels
        for i, m in enumerate(self.ch_mult):
            cfg.ch_mult[i] = m
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.pbe_vae_create(ctypes.byref(cfg), ctypes.byref(handle)), "pbe_vae_create")
            try:
                for name, p in self.state_dict().items():
                    host = p.detach().to("cpu", torch.float32).contiguous()
                    shape = (ctypes.c_int64 * host.dim())(*host.shape)
                    _lib.check(lib.pbe_vae_load_weight(handle, name.encode(), host.data_ptr(), shape, host.dim()),
                               f"pbe_vae_load_weight({name})")
                _lib.check(lib.pbe_vae_finalize_weights(handle), "pbe_vae_finalize_weights")
            except Exception:
                lib.pbe_vae_destroy(handle)
                raise
        self._engine, self._engine_version, self._engine_device = handle, self._weights_epoch, device

    # ------------------------------------------------------------------------------------------------------------
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """dec = Decoder(post_quant_conv(z)) (autoencoder.py:66-69). z: [B, embed_dim, h, w] CUDA -> [B, out_ch, 8h, 8w]."""
        if z.dim() != 4 or z.shape[1] != self.embed_dim:
            raise ValueError(f"expected z of shape [B,{self.embed_dim},h,w], got {tuple(z.shape)}")
        self._ensure_engine(z.device)
        zz = z.detach().to(torch.float32).contiguous()
        B, _, h, w = zz.shape
        f = 2 ** (len(self.ch_mult) - 1)
        out = torch.empty((B, self.out_ch, h * f, w * f), device=z.device, dtype=torch.float32)
        st = torch.cuda.current_stream(z.device).cuda_stream
        with torch.cuda.device(z.device):
            _lib.check(_lib.load().pbe_vae_decode(self._engine, zz.data_ptr(), out.data_ptr(), B, h, w, st), "pbe_vae_decode")
        return out.to(z.dtype)

    def decode_to_uint8(self, z: torch.Tensor) -> torch.Tensor:
        """decode(z) followed by the scripts' post-processing (scripts/inference.py:346-348,379-380) on the device:
        uint8 [B, H, W, 3] = trunc(255 * clamp((img + 1) / 2, 0, 1)), ready for PIL / PNG encoding on the host."""
        img = self.decode(z).to(torch.float32).contiguous()
        B, C, H, W = img.shape
        out = torch.empty((B, H, W, C), device=img.device, dtype=torch.uint8)
        st = torch.cuda.current_stream(img.device).cuda_stream
        with torch.cuda.device(img.device):
            _lib.check(_lib.load().pbe_postprocess_u8(img.data_ptr(), out.data_ptr(), B, C, H, W, st), "pbe_postprocess_u8")
        return out

    def encode(self, x: torch.Tensor) -> DiagonalGaussianDistribution:
        """posterior = DiagonalGaussianDistribution(quant_conv(Encoder(x))) (autoencoder.py:56-64).
        x: [B, in_channels, H, W] CUDA, H and W multiples of 8 * 2^(levels-1)."""
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise ValueError(f"expected x of shape [B,{self.in_channels},H,W], got {tuple(x.shape)}")
        self._ensure_engine(x.device)
        xx = x.detach().to(torch.float32).contiguous()
        B, _, H, W = xx.shape
        f = 2 ** (len(self.ch_mult) - 1)
        if H % f or W % f:
            raise ValueError(f"image height and width must be multiples of {f}")
        moments = torch.empty((B, 2 * self.embed_dim, H // f, W // f), device=x.device, dtype=torch.float32)
        st = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().pbe_vae_encode(self._engine, xx.data_ptr(), moments.data_ptr(), B, H, W, st),
                       "pbe_vae_encode")
        return DiagonalGaussianDistribution(moments.to(x.dtype))

    def forward(self, z):
        return self.decode(z)

    def profile(self, z: torch.Tensor, encode: bool = False):
        """Per-op device times of one eager decode (or encode): list of dicts(name, family, ms, flops)."""
        self._ensure_engine(z.device)
        lib = _lib.load()
        zz = z.detach().to(torch.float32).contiguous()
        B, _, h, w = zz.shape
        f = 2 ** (len(self.ch_mult) - 1)
        if encode:
            out = torch.empty((B, 2 * self.embed_dim, h // f, w // f), device=z.device, dtype=torch.float32)
        else:
            out = torch.empty((B, self.out_ch, h * f, w * f), device=z.device, dtype=torch.float32)
        ms = (ctypes.c_float * 4096)()
        st = torch.cuda.current_stream(z.device).cuda_stream
        with torch.cuda.device(z.device):
            n = lib.pbe_vae_profile(self._engine, int(encode), zz.data_ptr(), out.data_ptr(), B, h, w, st, ms, 4096)
        if n < 0:
            _lib.check(n, "pbe_vae_profile")
        rows = []
        for i in range(n):
            name, fam, fl = ctypes.c_char_p(), ctypes.c_char_p(), ctypes.c_double()
            _lib.check(lib.pbe_vae_op_info(self._engine, i, ctypes.byref(name), ctypes.byref(fam), ctypes.byref(fl)),
                       "pbe_vae_op_info")
            rows.append(dict(name=name.value.decode(), family=fam.value.decode(), ms=float(ms[i]), flops=fl.value))
        return rows
