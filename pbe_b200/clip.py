"""Drop-in for ``ldm.modules.encoders.modules.FrozenCLIPImageEmbedder`` (reference modules.py:138-171) — the second
"next" row of SURVEY.md §8f: the exemplar image becomes the single conditioning token of the U-Net,
``z = final_ln(mapper(CLIPVisionModel(image).pooler_output.unsqueeze(1)))``.

The reference constructor downloads ``openai/clip-vit-large-patch14``; this class is built from the architecture
numbers (defaults = ViT-L/14) and filled with ``load_state_dict`` — the state-dict keys are the reference's
(``transformer.vision_model.*`` as in ``transformers.CLIPVisionModel``, ``mapper.resblocks.*``, ``final_ln.*``), so a
Paint-by-Example checkpoint's ``cond_stage_model.*`` entries load unchanged.  ``forward`` / ``encode`` are one call into
the sm_100a library (``pbe_clip_encode`` in include/pbe_b200.h); there is no PyTorch/CPU fallback.
"""
from __future__ import annotations

import ctypes
from typing import Dict

import torch
import torch.nn as nn

from . import _lib
from .unet import _Node


class PbeClipConfig(ctypes.Structure):
    _fields_ = [("image_size", ctypes.c_int32), ("patch_size", ctypes.c_int32), ("width", ctypes.c_int32),
                ("layers", ctypes.c_int32), ("heads", ctypes.c_int32), ("mlp_dim", ctypes.c_int32),
                ("mapper_layers", ctypes.c_int32)]


def clip_param_shapes(image_size, patch_size, width, layers, heads, mlp_dim, mapper_layers) -> Dict[str, tuple]:
    """State-dict keys and shapes of FrozenCLIPImageEmbedder: transformers CLIPVisionModel (CLIPVisionTransformer) +
    xf.Transformer(1, width, mapper_layers, 1) + final_ln (modules.py:141-151, xf.py:107-130)."""
    s: Dict[str, tuple] = {}
    C, P = width, (image_size // patch_size) ** 2
    vm = "transformer.vision_model."

    def lin(p, o, i):
        s[p + ".weight"] = (o, i)
        s[p + ".bias"] = (o,)

    def norm(p):
        s[p + ".weight"] = (C,)
        s[p + ".bias"] = (C,)

    s[vm + "embeddings.class_embedding"] = (C,)
    s[vm + "embeddings.patch_embedding.weight"] = (C, 3, patch_size, patch_size)
    s[vm + "embeddings.position_embedding.weight"] = (P + 1, C)
    norm(vm + "pre_layrnorm")
    for i in range(layers):
        lp = f"{vm}encoder.layers.{i}."
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            lin(lp + "self_attn." + n, C, C)
        norm(lp + "layer_norm1")
        lin(lp + "mlp.fc1", mlp_dim, C)
        lin(lp + "mlp.fc2", C, mlp_dim)
        norm(lp + "layer_norm2")
    norm(vm + "post_layernorm")
    norm("final_ln")
    for j in range(mapper_layers):
        mp = f"mapper.resblocks.{j}."
        lin(mp + "attn.c_qkv", 3 * C, C)
        lin(mp + "attn.c_proj", C, C)
        norm(mp + "ln_1")
        lin(mp + "mlp.c_fc", 4 * C, C)
        lin(mp + "mlp.c_proj", C, 4 * C)
        norm(mp + "ln_2")
    return s


class FrozenCLIPImageEmbedder(nn.Module):
    def __init__(self, version="openai/clip-vit-large-patch14", image_size=224, patch_size=14, width=1024, layers=24,
                 heads=16, mlp_dim=4096, mapper_layers=5):
        super().__init__()
        self.version = version
        self.image_size, self.patch_size, self.width = int(image_size), int(patch_size), int(width)
        self.layers, self.heads, self.mlp_dim, self.mapper_layers = int(layers), int(heads), int(mlp_dim), int(mapper_layers)
        self._shapes = clip_param_shapes(self.image_size, self.patch_size, self.width, self.layers, self.heads,
                                         self.mlp_dim, self.mapper_layers)
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

    def freeze(self):
        return self

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
            _lib.load().pbe_clip_destroy(self._engine)
            self._engine = None

    def __del__(self):
        try:
            self._destroy_engine()
        except Exception:
            pass

    def _ensure_engine(self, device: torch.device):
        if device.type != "cuda":
            raise RuntimeError("pbe_b200.FrozenCLIPImageEmbedder runs only on a CUDA (sm_100a) device: no CPU fallback exists")
        if self._engine is not None and self._engine_version == self._weights_epoch and self._engine_device == device:
            return
        self._destroy_engine()
        lib = _lib.load()
        cfg = PbeClipConfig(self.image_size, self.patch_size, self.width, self.layers, self.heads, self.mlp_dim,
                            self.mapper_layers)
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.pbe_clip_create(ctypes.byref(cfg), ctypes.byref(handle)), "pbe_clip_create")
            try:
                for name, p in self.state_dict().items():
                    host = p.detach().to("cpu", torch.float32).contiguous()
                    shape = (ctypes.c_int64 * host.dim())(*host.shape)
                    _lib.check(lib.pbe_clip_load_weight(handle, name.encode(), host.data_ptr(), shape, host.dim()),
                               f"pbe_clip_load_weight({name})")
                _lib.check(lib.pbe_clip_finalize_weights(handle), "pbe_clip_finalize_weights")
            except Exception:
                lib.pbe_clip_destroy(handle)
                raise
        self._engine, self._engine_version, self._engine_device = handle, self._weights_epoch, device

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        """image [B, 3, S, S] (CLIP-normalised, CUDA) -> z [B, 1, width] (modules.py:160-166)."""
        if image.dim() != 4 or image.shape[1] != 3 or image.shape[2] != self.image_size or image.shape[3] != self.image_size:
            raise ValueError(f"expected image of shape [B,3,{self.image_size},{self.image_size}], got {tuple(image.shape)}")
        self._ensure_engine(image.device)
        x = image.detach().to(torch.float32).contiguous()
        B = x.shape[0]
        z = torch.empty((B, 1, self.width), device=image.device, dtype=torch.float32)
        st = torch.cuda.current_stream(image.device).cuda_stream
        with torch.cuda.device(image.device):
            _lib.check(_lib.load().pbe_clip_encode(self._engine, x.data_ptr(), z.data_ptr(), B, st), "pbe_clip_encode")
        return z.to(image.dtype)

    def encode(self, image):
        return self(image)
