"""ORACLE — test infrastructure only (never imported by the product path ``pbe_b200/``).

Plain-PyTorch fp32 restatement of the reference U-Net forward, written functionally over a state dict with the
reference's own key names.  Each function cites the reference code it restates (paths relative to /root/reference).
Pinned against the real reference modules by ``tests/test_oracle_pinned.py`` (live import, this container only) and
against the committed golden vectors in ``tests/golden/`` (generated from the real reference by
``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import torch
import torch.nn.functional as F

V1_CFG = dict(  # configs/v1.yaml:30-46
    in_channels=9, out_channels=4, model_channels=320, num_res_blocks=2, channel_mult=(1, 2, 4, 4),
    attention_resolutions=(4, 2, 1), num_heads=8, context_dim=768)

SMALL_CFG = dict(  # same topology, 5x narrower: CPU-fast parity config
    in_channels=9, out_channels=4, model_channels=64, num_res_blocks=2, channel_mult=(1, 2, 4, 4),
    attention_resolutions=(4, 2, 1), num_heads=8, context_dim=768)


def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    """ldm/modules/diffusionmodules/util.py:151-171."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half).to(
        device=timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def module_plan(cfg) -> list:
    """Execution order of UNetModel (openaimodel.py:558-834, forward :852-889) as
    (kind, state-dict prefix, cin, cout) tuples; kinds: conv_in, res, st, down, up, out, plus 'push'/'pop' markers
    encoded in the tuple's last field."""
    mc = cfg["model_channels"]
    plan = [("conv_in", "input_blocks.0.0", cfg["in_channels"], mc, "push")]
    chans = [mc]
    ch, ds, ib = mc, 1, 1
    nlev = len(cfg["channel_mult"])
    for level, mult in enumerate(cfg["channel_mult"]):
        for _ in range(cfg["num_res_blocks"]):
            attn = ds in cfg["attention_resolutions"]
            plan.append(("res", f"input_blocks.{ib}.0", ch, mult * mc, "" if attn else "push"))
            ch = mult * mc
            if attn:
                plan.append(("st", f"input_blocks.{ib}.1", ch, ch, "push"))
            chans.append(ch)
            ib += 1
        if level != nlev - 1:
            plan.append(("down", f"input_blocks.{ib}.0.op", ch, ch, "push"))
            chans.append(ch)
            ib += 1
            ds *= 2
    plan.append(("res", "middle_block.0", ch, ch, ""))
    plan.append(("st", "middle_block.1", ch, ch, ""))
    plan.append(("res", "middle_block.2", ch, ch, ""))
    ob = 0
    for level, mult in list(enumerate(cfg["channel_mult"]))[::-1]:
        for i in range(cfg["num_res_blocks"] + 1):
            ich = chans.pop()
            plan.append(("res", f"output_blocks.{ob}.0", ch + ich, mc * mult, "pop"))
            ch = mc * mult
            sub = 1
            if ds in cfg["attention_resolutions"]:
                plan.append(("st", f"output_blocks.{ob}.1", ch, ch, ""))
                sub = 2
            if level and i == cfg["num_res_blocks"]:
                plan.append(("up", f"output_blocks.{ob}.{sub}.conv", ch, ch, ""))
                ds //= 2
            ob += 1
    plan.append(("out", "out", ch, cfg["out_channels"], ""))
    return plan


def param_shapes(cfg) -> Dict[str, tuple]:
    """Every U-Net state-dict key and its shape (686 tensors for v1.yaml)."""
    mc, ctx = cfg["model_channels"], cfg["context_dim"]
    ted = 4 * mc
    s: Dict[str, tuple] = {}
    s["time_embed.0.weight"] = (ted, mc); s["time_embed.0.bias"] = (ted,)
    s["time_embed.2.weight"] = (ted, ted); s["time_embed.2.bias"] = (ted,)
    for kind, p, cin, cout, _ in module_plan(cfg):
        if kind in ("conv_in", "down", "up"):
            s[p + ".weight"] = (cout, cin, 3, 3); s[p + ".bias"] = (cout,)
        elif kind == "res":
            s[p + ".in_layers.0.weight"] = (cin,); s[p + ".in_layers.0.bias"] = (cin,)
            s[p + ".in_layers.2.weight"] = (cout, cin, 3, 3); s[p + ".in_layers.2.bias"] = (cout,)
            s[p + ".emb_layers.1.weight"] = (cout, ted); s[p + ".emb_layers.1.bias"] = (cout,)
            s[p + ".out_layers.0.weight"] = (cout,); s[p + ".out_layers.0.bias"] = (cout,)
            s[p + ".out_layers.3.weight"] = (cout, cout, 3, 3); s[p + ".out_layers.3.bias"] = (cout,)
            if cin != cout:
                s[p + ".skip_connection.weight"] = (cout, cin, 1, 1); s[p + ".skip_connection.bias"] = (cout,)
        elif kind == "st":
            c = cin
            s[p + ".norm.weight"] = (c,); s[p + ".norm.bias"] = (c,)
            s[p + ".proj_in.weight"] = (c, c, 1, 1); s[p + ".proj_in.bias"] = (c,)
            tb = p + ".transformer_blocks.0"
            for a, kdim in (("attn1", c), ("attn2", ctx)):
                s[f"{tb}.{a}.to_q.weight"] = (c, c)
                s[f"{tb}.{a}.to_k.weight"] = (c, kdim)
                s[f"{tb}.{a}.to_v.weight"] = (c, kdim)
                s[f"{tb}.{a}.to_out.0.weight"] = (c, c); s[f"{tb}.{a}.to_out.0.bias"] = (c,)
            s[tb + ".ff.net.0.proj.weight"] = (8 * c, c); s[tb + ".ff.net.0.proj.bias"] = (8 * c,)
            s[tb + ".ff.net.2.weight"] = (c, 4 * c); s[tb + ".ff.net.2.bias"] = (c,)
            for n in ("norm1", "norm2", "norm3"):
                s[f"{tb}.{n}.weight"] = (c,); s[f"{tb}.{n}.bias"] = (c,)
            s[p + ".proj_out.weight"] = (c, c, 1, 1); s[p + ".proj_out.bias"] = (c,)
        elif kind == "out":
            s["out.0.weight"] = (cin,); s["out.0.bias"] = (cin,)
            s["out.2.weight"] = (cout, cfg["model_channels"], 3, 3); s["out.2.bias"] = (cout,)
    return s


def res_block(sd, p: str, x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    """ResBlock._forward, openaimodel.py:255-275 (no up/down, no scale-shift norm); GroupNorm32 eps 1e-5 util.py:214."""
    h = F.group_norm(x.float(), 32, sd[p + ".in_layers.0.weight"], sd[p + ".in_layers.0.bias"], eps=1e-5)
    h = F.silu(h)
    h = F.conv2d(h, sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    emb_out = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
    h = h + emb_out[:, :, None, None]
    h = F.group_norm(h, 32, sd[p + ".out_layers.0.weight"], sd[p + ".out_layers.0.bias"], eps=1e-5)
    h = F.silu(h)
    h = F.conv2d(h, sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if (p + ".skip_connection.weight") in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


def cross_attention(sd, p: str, x: torch.Tensor, context, heads: int) -> torch.Tensor:
    """CrossAttention.forward, ldm/modules/attention.py:207-230 (literal: q, k, v, sim, softmax, out)."""
    q = F.linear(x, sd[p + ".to_q.weight"])
    ctx = x if context is None else context
    k = F.linear(ctx, sd[p + ".to_k.weight"])
    v = F.linear(ctx, sd[p + ".to_v.weight"])
    b, n, c = q.shape
    d = c // heads

    def split(t):
        return t.view(b, t.shape[1], heads, d).permute(0, 2, 1, 3).reshape(b * heads, t.shape[1], d)

    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bid,bjd->bij", q, k) * (d ** -0.5)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bij,bjd->bid", attn, v)
    out = out.view(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, c)
    return F.linear(out, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])


def spatial_transformer(sd, p: str, x: torch.Tensor, context: torch.Tensor, heads: int) -> torch.Tensor:
    """SpatialTransformer.forward attention.py:287-298 + BasicTransformerBlock._forward :248-252 + GEGLU :38-45."""
    b, c, hh, ww = x.shape
    x_in = x
    h = F.group_norm(x, 32, sd[p + ".norm.weight"], sd[p + ".norm.bias"], eps=1e-6)
    h = F.conv2d(h, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
    tb = p + ".transformer_blocks.0"
    h = cross_attention(sd, tb + ".attn1", F.layer_norm(h, (c,), sd[tb + ".norm1.weight"], sd[tb + ".norm1.bias"]),
                        None, heads) + h
    h = cross_attention(sd, tb + ".attn2", F.layer_norm(h, (c,), sd[tb + ".norm2.weight"], sd[tb + ".norm2.bias"]),
                        context, heads) + h
    n3 = F.layer_norm(h, (c,), sd[tb + ".norm3.weight"], sd[tb + ".norm3.bias"])
    proj = F.linear(n3, sd[tb + ".ff.net.0.proj.weight"], sd[tb + ".ff.net.0.proj.bias"])
    val, gate = proj.chunk(2, dim=-1)
    ff = F.linear(val * F.gelu(gate), sd[tb + ".ff.net.2.weight"], sd[tb + ".ff.net.2.bias"])
    h = ff + h
    h = h.reshape(b, hh, ww, c).permute(0, 3, 1, 2)
    h = F.conv2d(h, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return h + x_in


@torch.no_grad()
def unet_forward(sd: Dict[str, torch.Tensor], cfg, x: torch.Tensor, timesteps: torch.Tensor,
                 context: torch.Tensor, taps: dict | None = None) -> torch.Tensor:
    """UNetModel.forward, openaimodel.py:852-889. ``taps`` (optional dict) receives every module output by prefix."""
    mc = cfg["model_channels"]
    t_emb = timestep_embedding(timesteps, mc)
    emb = F.linear(t_emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    hs = []
    h = x.float()
    for kind, p, cin, cout, mark in module_plan(cfg):
        if mark == "pop":
            h = torch.cat([h, hs.pop()], dim=1)
        if kind == "conv_in":
            h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
        elif kind == "res":
            h = res_block(sd, p, h, emb)
        elif kind == "st":
            h = spatial_transformer(sd, p, h, context, cfg["num_heads"])
        elif kind == "down":
            h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], stride=2, padding=1)  # Downsample :150-160
        elif kind == "up":
            h = F.interpolate(h, scale_factor=2, mode="nearest")                      # Upsample :109-119
            h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
        elif kind == "out":
            h = F.group_norm(h, 32, sd["out.0.weight"], sd["out.0.bias"], eps=1e-5)
            h = F.conv2d(F.silu(h), sd["out.2.weight"], sd["out.2.bias"], padding=1)
        if taps is not None:
            taps[p] = h
        if mark == "push":
            hs.append(h)
    assert not hs
    return h


def make_state_dict(cfg, seed: int = 321, device="cpu") -> Dict[str, torch.Tensor]:
    """Deterministic, name-keyed random weights (no zero_module outputs: SURVEY.md §0.1 #8 — a freshly constructed
    reference U-Net outputs eps == 0, which would make parity vacuous). Independent of construction order, so the
    same function feeds the real reference module (load_state_dict), this oracle and the CUDA engine."""
    import zlib
    sd = {}
    for name, shape in param_shapes(cfg).items():
        g = torch.Generator(device="cpu").manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
        leaf = name.rsplit(".", 1)[1]
        is_norm = (len(shape) == 1 and leaf == "weight")
        if is_norm:
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        sd[name] = t.to(device)
    return sd
