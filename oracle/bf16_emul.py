"""ORACLE — test infrastructure only.  The fp32 oracle with the product's operand-rounding points emulated:
every conv / Linear sees 16-bit-rounded activations and weights (fp16, the library's default operand format, or bf16 --
``fmt``), every attention matmul bf16-rounded Q / K / P / V (the flash kernels stay bf16), all accumulating in fp32: the
arithmetic contract of the tcgen05 kernels.  The timestep-embedding MLP and the folded single-token
cross-attention stay fp32, as in the engine.  Comparing the CUDA path with THIS isolates kernel bugs from operand rounding
noise (which a narrow test network amplifies: with bf16 operands beyond the 1e-2 budget of the real v1.yaml width)."""
from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F

from . import unet_ref as M


def _rb(x):
    return x.bfloat16().float()


def _rh(x):
    return x.half().float()      # |x| > 65504 would saturate in the product; the test inputs stay far below


@contextlib.contextmanager
def _patched(fmt="f16"):
    of_conv, of_lin, of_einsum = F.conv2d, F.linear, torch.einsum
    _r = _rh if fmt == "f16" else _rb

    def conv2d(x, w, b=None, **kw):
        return of_conv(_r(x), _r(w), b, **kw)

    def linear(x, w, b=None):
        if x.dim() == 2:                      # time_embed / emb_layers: fp32 CUDA-core GEMV in the engine
            return of_lin(x, w, b)
        if x.dim() == 3 and x.shape[1] == 1:  # single-token context: folded in fp32
            return of_lin(x, w, b)
        return of_lin(_r(x), _r(w), b)

    def einsum(eq, a, b):
        return of_einsum(eq, _rb(a), _rb(b))

    M.F.conv2d, M.F.linear, M.torch.einsum = conv2d, linear, einsum
    try:
        yield
    finally:
        M.F.conv2d, M.F.linear, M.torch.einsum = of_conv, of_lin, of_einsum


@torch.no_grad()
def unet_forward_bf16(sd, cfg, x, t, context, fmt="f16"):
    """fmt: "f16" (the library default) or "bf16" (PBE_OPERANDS=bf16) GEMM operands; attention operands are bf16 in both."""
    with _patched(fmt):
        return M.unet_forward(sd, cfg, x, t, context)
