"""Drop-in for ``ldm.modules.diffusionmodules.openaimodel.UNetModel`` (reference openaimodel.py:528-889).

Same constructor keywords (the subset ``configs/v1.yaml:30-46`` uses), same state-dict keys (so
``load_state_dict`` of a Paint-by-Example checkpoint under ``model.diffusion_model.*`` fills it), same
``forward(x, timesteps, context)`` contract — but the forward is one call into the sm_100a engine
(``pbe_unet_forward`` in include/pbe_b200.h).  There is no PyTorch/CPU fallback: without a CUDA device or the
compiled library the forward raises.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib


class PbeConfig(ctypes.Structure):
    _fields_ = [("in_channels", ctypes.c_int32), ("out_channels", ctypes.c_int32),
                ("model_channels", ctypes.c_int32), ("num_res_blocks", ctypes.c_int32),
                ("num_levels", ctypes.c_int32), ("channel_mult", ctypes.c_int32 * 8),
                ("num_attention_resolutions", ctypes.c_int32), ("attention_resolutions", ctypes.c_int32 * 8),
                ("num_heads", ctypes.c_int32), ("context_dim", ctypes.c_int32)]


def unet_param_shapes(in_channels, out_channels, model_channels, num_res_blocks, channel_mult,
                      attention_resolutions, num_heads, context_dim) -> Dict[str, tuple]:
    """State-dict keys and shapes of the reference U-Net for use_spatial_transformer=True, transformer_depth=1,
    resblock_updown=False (mirrors UNetModel.__init__, openaimodel.py:558-834)."""
    mc, ctx, ted = model_channels, context_dim, 4 * model_channels
    s: Dict[str, tuple] = {}

    def lin(p, o, i, bias=True):
        s[p + ".weight"] = (o, i)
        if bias:
            s[p + ".bias"] = (o,)

    def conv(p, o, i, k):
        s[p + ".weight"] = (o, i, k, k)
        s[p + ".bias"] = (o,)

    def norm(p, c):
        s[p + ".weight"] = (c,)
        s[p + ".bias"] = (c,)

    def res(p, cin, cout):
        norm(p + ".in_layers.0", cin); conv(p + ".in_layers.2", cout, cin, 3)
        lin(p + ".emb_layers.1", cout, ted)
        norm(p + ".out_layers.0", cout); conv(p + ".out_layers.3", cout, cout, 3)
        if cin != cout:
            conv(p + ".skip_connection", cout, cin, 1)

    def st(p, c):
        norm(p + ".norm", c); conv(p + ".proj_in", c, c, 1)
        tb = p + ".transformer_blocks.0"
        for a, kd in (("attn1", c), ("attn2", ctx)):
            lin(f"{tb}.{a}.to_q", c, c, False); lin(f"{tb}.{a}.to_k", c, kd, False); lin(f"{tb}.{a}.to_v", c, kd, False)
            lin(f"{tb}.{a}.to_out.0", c, c)
        lin(tb + ".ff.net.0.proj", 8 * c, c); lin(tb + ".ff.net.2", c, 4 * c)
        for n in ("norm1", "norm2", "norm3"):
            norm(f"{tb}.{n}", c)
        conv(p + ".proj_out", c, c, 1)

    lin("time_embed.0", ted, mc); lin("time_embed.2", ted, ted)
    conv("input_blocks.0.0", mc, in_channels, 3)
    chans = [mc]
    ch, ds, ib = mc, 1, 1
    for level, mult in enumerate(channel_mult):
        for _ in range(num_res_blocks):
            res(f"input_blocks.{ib}.0", ch, mult * mc)
            ch = mult * mc
            if ds in attention_resolutions:
                st(f"input_blocks.{ib}.1", ch)
            chans.append(ch)
            ib += 1
        if level != len(channel_mult) - 1:
            conv(f"input_blocks.{ib}.0.op", ch, ch, 3)
            chans.append(ch)
            ib += 1
            ds *= 2
    res("middle_block.0", ch, ch); st("middle_block.1", ch); res("middle_block.2", ch, ch)
    ob = 0
    for level, mult in list(enumerate(channel_mult))[::-1]:
        for i in range(num_res_blocks + 1):
            ich = chans.pop()
            res(f"output_blocks.{ob}.0", ch + ich, mc * mult)
            ch = mc * mult
            sub = 1
            if ds in attention_resolutions:
                st(f"output_blocks.{ob}.1", ch)
                sub = 2
            if level and i == num_res_blocks:
                conv(f"output_blocks.{ob}.{sub}.conv", ch, ch, 3)
                ds //= 2
            ob += 1
    norm("out.0", ch); conv("out.2", out_channels, mc, 3)
    return s


class _Node(nn.Module):
    """Anonymous container: gives parameters the reference's dotted names."""


class UNetModel(nn.Module):
    def __init__(self, image_size=32, in_channels=9, model_channels=320, out_channels=4, num_res_blocks=2,
                 attention_resolutions=(4, 2, 1), dropout=0, channel_mult=(1, 2, 4, 4), conv_resample=True, dims=2,
                 num_classes=None, use_checkpoint=False, use_fp16=False, num_heads=8, num_head_channels=-1,
                 num_heads_upsample=-1, use_scale_shift_norm=False, resblock_updown=False,
                 use_new_attention_order=False, use_spatial_transformer=True, transformer_depth=1, context_dim=768,
                 n_embed=None, legacy=False, add_conv_in_front_of_unet=False):
        super().__init__()
        unsupported = dict(dims=(dims, 2), num_classes=(num_classes, None), use_scale_shift_norm=(use_scale_shift_norm, False),
                           resblock_updown=(resblock_updown, False), use_spatial_transformer=(use_spatial_transformer, True),
                           transformer_depth=(transformer_depth, 1), n_embed=(n_embed, None),
                           add_conv_in_front_of_unet=(add_conv_in_front_of_unet, False), conv_resample=(conv_resample, True),
                           num_head_channels=(num_head_channels, -1), use_fp16=(use_fp16, False))
        for k, (v, want) in unsupported.items():
            if v != want:
                raise NotImplementedError(f"pbe_b200.UNetModel supports {k}={want!r} only (configs/v1.yaml), got {v!r}")
        if dropout:
            raise NotImplementedError("inference only: dropout must be 0")
        if isinstance(context_dim, (list, tuple)) or type(context_dim).__name__ == "ListConfig":
            context_dim = list(context_dim)[0]
        self.image_size = image_size
        self.in_channels = int(in_channels)
        self.out_channels = int(out_channels)
        self.model_channels = int(model_channels)
        self.num_res_blocks = int(num_res_blocks)
        self.attention_resolutions = tuple(int(a) for a in attention_resolutions)
        self.channel_mult = tuple(int(m) for m in channel_mult)
        self.num_heads = int(num_heads)
        self.context_dim = int(context_dim)
        self.dtype = torch.float32
        self._shapes = unet_param_shapes(self.in_channels, self.out_channels, self.model_channels,
                                         self.num_res_blocks, self.channel_mult, self.attention_resolutions,
                                         self.num_heads, self.context_dim)
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
        self._ctx_key = None

    # ------------------------------------------------------------------------------------------------------------
    def mark_weights_changed(self) -> None:
        """Call after mutating parameters in place; load_state_dict / .to() / .cuda() are tracked automatically."""
        self._weights_epoch += 1

    def _load_from_state_dict(self, *args, **kwargs):
        self._weights_epoch += 1
        return super()._load_from_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._weights_epoch += 1
        return super()._apply(fn, *args, **kwargs)

    def _weights_version(self):
        return self._weights_epoch

    def _destroy_engine(self):
        if self._engine is not None:
            _lib.load().pbe_destroy(self._engine)
            self._engine = None

    def __del__(self):
        try:
            self._destroy_engine()
        except Exception:
            pass

    def _ensure_engine(self, device: torch.device):
        if device.type != "cuda":
            raise RuntimeError("pbe_b200.UNetModel runs only on a CUDA (sm_100a) device: no CPU fallback exists")
        ver = self._weights_version()
        if self._engine is not None and self._engine_version == ver and self._engine_device == device:
            return
        self._destroy_engine()
        lib = _lib.load()
        cfg = PbeConfig()
        cfg.in_channels, cfg.out_channels, cfg.model_channels = self.in_channels, self.out_channels, self.model_channels
        cfg.num_res_blocks, cfg.num_levels = self.num_res_blocks, len(self.channel_mult)
        for i, m in enumerate(self.channel_mult):
            cfg.channel_mult[i] = m
        cfg.num_attention_resolutions = len(self.attention_resolutions)
        for i, a in enumerate(self.attention_resolutions):
            cfg.attention_resolutions[i] = a
        cfg.num_heads, cfg.context_dim = self.num_heads, self.context_dim
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.pbe_create(ctypes.byref(cfg), ctypes.byref(handle)), "pbe_create")
            try:
                for name, p in self.state_dict().items():
                    host = p.detach().to("cpu", torch.float32).contiguous()
                    shape = (ctypes.c_int64 * host.dim())(*host.shape)
                    _lib.check(lib.pbe_load_weight(handle, name.encode(), host.data_ptr(), shape, host.dim()),
                               f"pbe_load_weight({name})")
                _lib.check(lib.pbe_finalize_weights(handle), "pbe_finalize_weights")
            except Exception:
                lib.pbe_destroy(handle)
                raise
        self._engine, self._engine_version, self._engine_device = handle, ver, device
        self._ctx_key = None

    # ------------------------------------------------------------------------------------------------------------
    def set_context(self, context: torch.Tensor) -> None:
        """Fold the single-token cross-attention for ``context`` [Bc, 1, context_dim] (K4, attention.py:207-230)."""
        if context.dim() != 3 or context.shape[1] != 1 or context.shape[2] != self.context_dim:
            raise NotImplementedError(
                f"pbe_b200 folds the Paint-by-Example single-token conditioning [B,1,{self.context_dim}]; "
                f"got context of shape {tuple(context.shape)}")
        self._ensure_engine(context.device)
        ctx = context.detach().to(torch.float32).reshape(context.shape[0], self.context_dim).contiguous()
        st = torch.cuda.current_stream(ctx.device).cuda_stream
        with torch.cuda.device(ctx.device):
            _lib.check(_lib.load().pbe_set_context(self._engine, ctx.data_ptr(), ctx.shape[0], st), "pbe_set_context")
        self._ctx_batch = ctx.shape[0]
        self._ctx_version = getattr(self, "_ctx_version", 0) + 1     # samplers compare this to notice a context they did not set

    @property
    def context_version(self) -> int:
        """Incremented by every :meth:`set_context` (and ``forward``): a sampling loop that set its context remembers the
        value and re-applies its own context when a callback / another caller on the same stream has changed it since."""
        return getattr(self, "_ctx_version", 0)

    def run(self, x: torch.Tensor, timesteps: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """eps = UNet(x, t) with the context set by :meth:`set_context`. x: [Bc, in_ch, H, W] fp32 CUDA."""
        self._ensure_engine(x.device)
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise ValueError(f"expected x of shape [B,{self.in_channels},H,W], got {tuple(x.shape)}")
        x = x.detach().to(torch.float32).contiguous()
        t = timesteps.detach().to(device=x.device, dtype=torch.int64).contiguous()
        if t.shape[0] != x.shape[0]:
            raise ValueError("timesteps batch != x batch")
        Bc, _, H, W = x.shape
        if out is None:
            out = torch.empty((Bc, self.out_channels, H, W), device=x.device, dtype=torch.float32)
        st = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().pbe_unet_forward(self._engine, x.data_ptr(), t.data_ptr(), out.data_ptr(), Bc, H, W,
                                                    st), "pbe_unet_forward")
        return out

    def run_cfg_pair(self, x: torch.Tensor, timesteps: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The CFG evaluation ``apply_model(cat([x]*2), cat([t]*2), cat([uc, c]))`` (plms.py:185-188) with x [B, in_ch, H, W]
        and t [B] given once; the context (2B rows, unconditional first) comes from :meth:`set_context`.  Returns
        eps [2B, out_ch, H, W], bit-identical to :meth:`run` on the duplicated batch: the layers in front of the first
        cross-attention are evaluated once for the B shared samples (``pbe_unet_forward_cfg_pair``)."""
        self._ensure_engine(x.device)
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise ValueError(f"expected x of shape [B,{self.in_channels},H,W], got {tuple(x.shape)}")
        x = x.detach().to(torch.float32).contiguous()
        t = timesteps.detach().to(device=x.device, dtype=torch.int64).contiguous()
        if t.shape[0] != x.shape[0]:
            raise ValueError("timesteps batch != x batch")
        B, _, H, W = x.shape
        if out is None:
            out = torch.empty((2 * B, self.out_channels, H, W), device=x.device, dtype=torch.float32)
        st = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().pbe_unet_forward_cfg_pair(self._engine, x.data_ptr(), t.data_ptr(), out.data_ptr(), B,
                                                             H, W, st), "pbe_unet_forward_cfg_pair")
        return out

    def forward(self, x, timesteps=None, context=None, y=None, **kwargs):
        """Same contract as the reference ``UNetModel.forward`` (openaimodel.py:852-889)."""
        if y is not None:
            raise NotImplementedError("class-conditional U-Net is not part of the Paint-by-Example path")
        if context is None:
            raise ValueError("context (the exemplar embedding) is required")
        self.set_context(context)
        return self.run(x, timesteps).to(x.dtype)

    def profile(self, x: torch.Tensor, timesteps: torch.Tensor):
        """Per-op device times of one eager forward: list of dicts(name, family, ms, flops, bytes)."""
        self._ensure_engine(x.device)
        lib = _lib.load()
        x = x.detach().to(torch.float32).contiguous()
        t = timesteps.detach().to(device=x.device, dtype=torch.int64).contiguous()
        Bc, _, H, W = x.shape
        out = torch.empty((Bc, self.out_channels, H, W), device=x.device, dtype=torch.float32)
        ms = (ctypes.c_float * 4096)()
        st = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            n = lib.pbe_profile_forward(self._engine, x.data_ptr(), t.data_ptr(), out.data_ptr(), Bc, H, W, st, ms, 4096)
        if n < 0:
            _lib.check(n, "pbe_profile_forward")
        rows = []
        for i in range(n):
            name, fam = ctypes.c_char_p(), ctypes.c_char_p()
            fl, by = ctypes.c_double(), ctypes.c_double()
            _lib.check(lib.pbe_op_info(self._engine, i, ctypes.byref(name), ctypes.byref(fam), ctypes.byref(fl),
                                       ctypes.byref(by)), "pbe_op_info")
            rows.append(dict(name=name.value.decode(), family=fam.value.decode(), ms=float(ms[i]), flops=fl.value,
                             bytes=by.value))
        return rows

    def launches_per_forward(self) -> int:
        return int(_lib.load().pbe_launches_per_forward(self._engine)) if self._engine is not None else 0

    def set_use_graph(self, enable: bool) -> None:
        self._ensure_engine(next(self.parameters()).device)
        _lib.check(_lib.load().pbe_set_use_graph(self._engine, int(enable)), "pbe_set_use_graph")
