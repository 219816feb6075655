"""Drop-in for the decode half of ``ldm.models.autoencoder.AutoencoderKL`` (reference autoencoder.py:14-69) — the first
"next" row of SURVEY.md §8f: ``LatentDiffusion.decode_first_stage`` (latent_diffusion.py:444-508) turns the sampled
latents into images once per request, right after the PLMS/DDIM loop.

Same constructor keywords as the reference (``ddconfig``, ``embed_dim``; ``lossconfig`` & co. are accepted and ignored),
same state-dict keys for the decoder half (``decoder.*``, ``post_quant_conv.*`` — a Paint-by-Example checkpoint's
``first_stage_model.*`` entries load with ``strict=False``; the ``encoder.*`` / ``quant_conv.*`` / ``loss.*`` keys are not
used here), same ``decode(z)`` contract.  The decode is one call into the sm_100a library (``pbe_vae_decode`` in
include/pbe_b200.h); there is no PyTorch/CPU fallback.  ``encode`` is not part of this row and raises.
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
                ("num_res_blocks", ctypes.c_int32)]


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
        self.embed_dim = int(embed_dim)
        self.z_channels = int(dd["z_channels"])
        self.ch = int(dd["ch"])
        self.out_ch = int(dd["out_ch"])
        self.ch_mult = tuple(int(m) for m in dd["ch_mult"])
        self.num_res_blocks = int(dd["num_res_blocks"])
        self._shapes = vae_decoder_param_shapes(self.embed_dim, self.z_channels, self.ch, self.out_ch, self.ch_mult,
                                                self.num_res_blocks)
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
        cfg.num_levels, cfg.num_res_blocks = len(self.ch_mult), self.num_res_blocks
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

    def encode(self, x):
        raise NotImplementedError("VAE encode is outside this row (SURVEY.md §8f rank 1 covers it next); attach the "
                                  "reference encoder for it")

    def forward(self, z):
        return self.decode(z)

    def profile(self, z: torch.Tensor):
        """Per-op device times of one eager decode: list of dicts(name, family, ms, flops)."""
        self._ensure_engine(z.device)
        lib = _lib.load()
        zz = z.detach().to(torch.float32).contiguous()
        B, _, h, w = zz.shape
        f = 2 ** (len(self.ch_mult) - 1)
        out = torch.empty((B, self.out_ch, h * f, w * f), device=z.device, dtype=torch.float32)
        ms = (ctypes.c_float * 4096)()
        st = torch.cuda.current_stream(z.device).cuda_stream
        with torch.cuda.device(z.device):
            n = lib.pbe_vae_profile_decode(self._engine, zz.data_ptr(), out.data_ptr(), B, h, w, st, ms, 4096)
        if n < 0:
            _lib.check(n, "pbe_vae_profile_decode")
        rows = []
        for i in range(n):
            name, fam, fl = ctypes.c_char_p(), ctypes.c_char_p(), ctypes.c_double()
            _lib.check(lib.pbe_vae_op_info(self._engine, i, ctypes.byref(name), ctypes.byref(fam), ctypes.byref(fl)),
                       "pbe_vae_op_info")
            rows.append(dict(name=name.value.decode(), family=fam.value.decode(), ms=float(ms[i]), flops=fl.value))
        return rows
