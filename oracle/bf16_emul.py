"""ORACLE — test infrastructure only.  The fp32 oracle with the product's operand-rounding points emulated:
every conv / Linear / attention matmul sees bf16-rounded activations and weights and accumulates in fp32, exactly the
arithmetic contract of the tcgen05 kernels.  The timestep-embedding MLPs and the folded
single-token cross-attention stay fp32, as in the engine.  Comparing the CUDA path with THIS isolates kernel bugs from
bf16 rounding noise (which a narrow test network amplifies beyond the 1e-2 budget of the real v1.yaml width)."""
from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F

from . import unet_ref as M


def _r(x):
    return x.bfloat16().float()


@contextlib.contextmanager
def _patched():
    of_conv, of_lin, of_einsum = F.conv2d, F.linear, torch.einsum

    def conv2d(x, w, b=None, **kw):
        return of_conv(_r(x), _r(w), b, **kw)

    def linear(x, w, b=None):
        if x.dim() == 2:                      # time_embed / emb_layers: fp32 CUDA-core GEMV in the engine
            return of_lin(x, w, b)
        if x.dim() == 3 and x.shape[1] == 1:  # single-token context: folded in fp32
            return of_lin(x, w, b)
        return of_lin(_r(x), _r(w), b)

    def einsum(eq, a, b):
        return of_einsum(eq, _r(a), _r(b))

    M.F.conv2d, M.F.linear, M.torch.einsum = conv2d, linear, einsum
    try:
        yield
    finally:
        M.F.conv2d, M.F.linear, M.torch.einsum = of_conv, of_lin, of_einsum


@torch.no_grad()
def unet_forward_bf16(sd, cfg, x, t, context):
    with _patched():
        return M.unet_forward(sd, cfg, x, t, context)
