"""Parity of every CUDA operator (called through the C ABI) against torch fp32 restatements of the reference ops."""
import ctypes
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# The 16-bit GEMM operand format is a process-wide library setting (include/pbe_b200.h: pbe_set_operand_format): fp16 by
# default, bf16 on request.  Every test of this module runs in both; OP16 is the torch dtype of the current one.  (The
# self-attention entry point always READS bf16 Q | K | V^T -- its P matrix lives far below the fp16 range -- and writes its
# output in the operand format.)
_FMT = {"dtype": torch.float16}


def OP16():
    return _FMT["dtype"]


@pytest.fixture(scope="module", params=["f16", "bf16"])
def lib(request):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from pbe_b200 import _lib
    handle = _lib.load()
    f16 = request.param == "f16"
    assert handle.pbe_set_operand_format(1 if f16 else 0) == 0
    _FMT["dtype"] = torch.float16 if f16 else torch.bfloat16
    yield handle
    handle.pbe_set_operand_format(1)          # the library default
    _FMT["dtype"] = torch.float16


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def _err(lib):
    return lib.pbe_last_error().decode()


# ------------------------------------------------------------------------------------------------------------------
GEMM_CASES = [
    # Nb, H, W, C, k, stride, Cout, block_n, rowbias, residual
    (1, 1, 256, 64, 1, 1, 128, 128, False, False),
    (1, 1, 1000, 128, 1, 1, 128, 128, True, True),     # ragged M
    (1, 1, 512, 320, 1, 1, 320, 160, False, False),
    (1, 1, 300, 320, 1, 1, 4, 32, False, False),       # tiny N (final conv shape)
    (2, 64, 64, 64, 3, 1, 128, 0, False, False),
    (2, 32, 32, 128, 3, 1, 160, 0, True, True),
    (3, 8, 8, 128, 3, 1, 128, 0, False, True),          # tile spans samples
    (2, 2, 2, 64, 3, 1, 64, 0, False, False),           # smaller than a tile
    (1, 96, 96, 64, 3, 1, 128, 0, False, False),        # 768x768 latent width
    (3, 12, 12, 64, 3, 1, 128, 0, False, False),
    (2, 64, 64, 64, 3, 2, 128, 0, False, False),        # Downsample
    (2, 16, 16, 128, 3, 2, 128, 0, True, False),
    (2, 32, 32, 192, 1, 1, 128, 0, False, True),
]


@pytest.mark.parametrize("case", GEMM_CASES)
def test_conv_gemm_matches_conv2d(lib, case):
    Nb, H, W, C, k, stride, Cout, bn, use_rb, use_res = case
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1234 + sum(case[:7]))
    x = torch.randn(Nb, H, W, C, generator=g).to(dev).to(OP16())
    w = (torch.randn(Cout, C, k, k, generator=g) / math.sqrt(C * k * k)).to(dev).to(OP16())
    wt = w.permute(2, 3, 0, 1).contiguous().view(k * k, Cout, C)
    bias = torch.randn(Cout, generator=g).to(dev)
    Ho, Wo = H // stride, W // stride
    rb = torch.randn(Nb, Cout, generator=g).to(dev) if use_rb else None
    res = torch.randn(Nb, Ho, Wo, Cout, generator=g).to(dev) if use_res else None
    out = torch.full((Nb, Ho, Wo, Cout), float("nan"), device=dev)
    outb = torch.zeros((Nb, Ho, Wo, Cout), device=dev, dtype=OP16())
    rc = lib.pbe_op_conv_gemm(x.data_ptr(), Nb, H, W, C, k, stride, wt.data_ptr(), Cout, 0, _p(bias), _p(rb), _p(res),
                              out.data_ptr(), outb.data_ptr(), None, 0, bn, _stream())
    assert rc == 0, _err(lib)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, stride=stride, padding=k // 2).permute(0, 2, 3, 1)
    if rb is not None:
        ref = ref + rb[:, None, None, :]
    if res is not None:
        ref = ref + res
    assert not torch.isnan(out).any()
    assert _rel(out, ref) < 1e-5                      # same bf16 inputs, fp32 accumulation: only summation order differs
    assert _rel(outb, ref) < 5e-3                     # bf16 rounding of the output


@pytest.mark.parametrize("case", [(2, 64, 64, 64, 3, 320, True), (2, 32, 32, 128, 1, 160, False), (3, 16, 16, 64, 3, 128, True)])
def test_gemm_fused_groupnorm_statistics(lib, case):
    """The epilogue's fused GroupNorm statistics: (sum, sum of squares) of the fp32 output over each run of 32 rows."""
    Nb, H, W, C, k, Cout, use_res = case
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(77 + sum(case[:6]))
    x = torch.randn(Nb, H, W, C, generator=g).to(dev).to(OP16())
    w = (torch.randn(Cout, C, k, k, generator=g) / math.sqrt(C * k * k)).to(dev).to(OP16())
    wt = w.permute(2, 3, 0, 1).contiguous().view(k * k, Cout, C)
    bias = torch.randn(Cout, generator=g).to(dev)
    res = torch.randn(Nb, H, W, Cout, generator=g).to(dev) if use_res else None
    out = torch.zeros((Nb, H, W, Cout), device=dev)
    M = Nb * H * W
    stats = torch.full((M // 32, Cout, 2), float("nan"), device=dev)
    lib.pbe_debug_set_gemm_stats_out(stats.data_ptr())
    try:
        rc = lib.pbe_op_conv_gemm(x.data_ptr(), Nb, H, W, C, k, 1, wt.data_ptr(), Cout, 0, _p(bias), None, _p(res),
                                  out.data_ptr(), None, None, 0, 0, _stream())
    finally:
        lib.pbe_debug_set_gemm_stats_out(None)
    assert rc == 0, _err(lib)
    torch.cuda.synchronize()
    o = out.view(M // 32, 32, Cout).double()
    assert torch.allclose(stats[..., 0].double(), o.sum(1), rtol=1e-5, atol=1e-3)
    assert torch.allclose(stats[..., 1].double(), (o * o).sum(1), rtol=1e-5, atol=1e-3)


def test_gemm_geglu_epilogue(lib):
    """GEGLU: value * gelu_erf(gate), ldm/modules/attention.py:43-45, with the per-tile interleaved weight layout."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(7)
    M, C = 640, 128
    inner = 4 * C
    x = torch.randn(M, C, generator=g).to(dev).to(OP16())
    w = (torch.randn(2 * inner, C, generator=g) / math.sqrt(C)).to(dev).to(OP16())
    b = torch.randn(2 * inner, generator=g).to(dev)
    # interleave: tile t = 128 value rows then 128 gate rows (engine.cu add_st)
    wv, wg = w[:inner].view(inner // 128, 128, C), w[inner:].view(inner // 128, 128, C)
    wi = torch.cat((wv, wg), dim=1).reshape(2 * inner, C).contiguous()
    bi = torch.cat((b[:inner].view(-1, 128), b[inner:].view(-1, 128)), dim=1).reshape(-1).contiguous()
    out = torch.zeros(M, inner, device=dev, dtype=OP16())
    rc = lib.pbe_op_conv_gemm(x.data_ptr(), 1, 1, M, C, 1, 1, wi.data_ptr(), 2 * inner, 1, bi.data_ptr(), None, None,
                              None, out.data_ptr(), None, 0, 0, _stream())
    assert rc == 0, _err(lib)
    torch.cuda.synchronize()
    proj = F.linear(x.float(), w.float(), b)
    val, gate = proj.chunk(2, dim=-1)
    ref = val * F.gelu(gate)
    assert _rel(out, ref) < 5e-3


def test_gemm_qkv_split_store(lib):
    """Fused q|k|v projection: Q|K row-major, V transposed per sample."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(9)
    B, H, W, C = 2, 16, 16, 128
    N = H * W
    x = torch.randn(B, H, W, C, generator=g).to(dev).to(OP16())
    w = (torch.randn(3 * C, C, generator=g) / math.sqrt(C)).to(dev).to(OP16())
    qk = torch.zeros(B, N, 2 * C, device=dev, dtype=OP16())
    vt = torch.zeros(B, C, N, device=dev, dtype=OP16())
    rc = lib.pbe_op_conv_gemm(x.data_ptr(), B, H, W, C, 1, 1, w.data_ptr(), 3 * C, 2, None, None, None, None,
                              qk.data_ptr(), vt.data_ptr(), 2 * C, 128, _stream())
    assert rc == 0, _err(lib)
    torch.cuda.synchronize()
    ref = F.linear(x.float().view(B, N, C), w.float())
    assert _rel(qk, ref[..., :2 * C]) < 5e-3
    assert _rel(vt, ref[..., 2 * C:].transpose(1, 2)) < 5e-3


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,N,heads,d", [(2, 4096, 8, 40), (2, 1024, 8, 80), (2, 256, 8, 160), (2, 64, 8, 160),
                                         (1, 16, 8, 8), (1, 1024, 8, 16), (1, 200, 8, 32), (1, 2304, 8, 80),
                                         (1, 144, 8, 160), (2, 180, 8, 160), (1, 45, 8, 160), (2, 1035, 8, 40),
                                         (2, 257, 16, 64), (1, 26, 2, 64)])
def test_self_attention_matches_softmax_reference(lib, B, N, heads, d):
    """softmax(q k^T d^-1/2) v, ldm/modules/attention.py:217-229."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(100 + N + d)
    C = heads * d
    q = torch.randn(B, N, C, generator=g).to(dev).bfloat16()
    k = torch.randn(B, N, C, generator=g).to(dev).bfloat16()
    v = torch.randn(B, N, C, generator=g).to(dev).bfloat16()
    qk = torch.cat((q, k), dim=-1).contiguous()
    Np = (N + 7) // 8 * 8                       # V^T rows are pitched to a multiple of 8 tokens (pad never read)
    vt = torch.full((B, C, Np), float("nan"), device=dev, dtype=torch.bfloat16)
    vt[:, :, :N] = v.transpose(1, 2)
    out = torch.zeros(B, N, C, device=dev, dtype=OP16())
    rc = lib.pbe_op_self_attention(qk.data_ptr(), vt.data_ptr(), out.data_ptr(), B, N, heads, d, _stream())
    assert rc == 0, _err(lib)
    torch.cuda.synchronize()

    def split(t):
        return t.float().view(B, N, heads, d).permute(0, 2, 1, 3)

    sim = torch.einsum("bhid,bhjd->bhij", split(q), split(k)) * d ** -0.5
    ref = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), split(v)).permute(0, 2, 1, 3).reshape(B, N, C)
    assert not torch.isnan(out.float()).any()
    assert _rel(out, ref) < 1e-2, _rel(out, ref)


@pytest.mark.parametrize("N,d,heads,drift", [(4096, 40, 8, 12.0), (4096, 40, 8, 45.0), (1035, 40, 8, 40.0),
                                             (257, 64, 16, 30.0), (2048, 64, 4, -50.0), (4096, 40, 8, 3.0),
                                             (4096, 40, 8, 1.0), (1035, 40, 8, 1.5), (2048, 64, 4, 2.0),
                                             (1024, 80, 8, 2.0), (1024, 80, 8, 30.0), (2304, 80, 8, -40.0),
                                             (256, 160, 8, 3.0), (576, 160, 8, 25.0), (1024, 16, 8, 4.0)])
def test_self_attention_drifting_maximum(lib, N, d, heads, drift):
    """The single-pass tiles of flash_attn2_kernel (d <= 64) take their reference maximum two key tiles late, those of
    flash_attn_kernel (d = 80, 160) one tile late: logits that climb
    (or fall) by `drift` nats per 128 keys exercise the stale reference, the lazy rescale of O (small drifts: the old
    contributions still matter after a rescale, so a wrong rescale factor shows) and the headroom
    (ATT2_BIAS) without reaching the documented saturation bound (a logit 111 nats above every key of the tiles up to
    j-2, i.e. a sustained climb of > 55 nats per 128 keys).  Same softmax as attention.py:217-229."""
    dev = torch.device("cuda:0")
    B = 1
    g = torch.Generator().manual_seed(7 + N + d)
    C = heads * d
    q = torch.randn(B, N, heads, d, generator=g)
    k = torch.randn(B, N, heads, d, generator=g) * 0.5
    # one direction per head carries a ramp over the key index: q.k * d^-1/2 gains `drift` nats per 128 keys
    u = torch.nn.functional.normalize(torch.randn(heads, d, generator=g), dim=-1)
    q = q - (q * u).sum(-1, keepdim=True) * u + 4.0 * u           # every query has component 4 along u
    ramp = torch.arange(N, dtype=torch.float32) / 128.0 * drift * math.sqrt(d) / 4.0
    k = k - (k * u).sum(-1, keepdim=True) * u + ramp.view(1, N, 1, 1) * u
    q = q.reshape(B, N, C).to(dev).bfloat16()
    k = k.reshape(B, N, C).to(dev).bfloat16()
    v = torch.randn(B, N, C, generator=g).to(dev).bfloat16()
    qk = torch.cat((q, k), dim=-1).contiguous()
    Np = (N + 7) // 8 * 8
    vt = torch.zeros(B, C, Np, device=dev, dtype=torch.bfloat16)
    vt[:, :, :N] = v.transpose(1, 2)
    out = torch.zeros(B, N, C, device=dev, dtype=OP16())
    rc = lib.pbe_op_self_attention(qk.data_ptr(), vt.data_ptr(), out.data_ptr(), B, N, heads, d, _stream())
    assert rc == 0, _err(lib)
    torch.cuda.synchronize()

    def split(t):
        return t.float().view(B, N, heads, d).permute(0, 2, 1, 3)

    sim = torch.einsum("bhid,bhjd->bhij", split(q), split(k)) * d ** -0.5
    ref = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), split(v)).permute(0, 2, 1, 3).reshape(B, N, C)
    assert torch.isfinite(out.float()).all()
    assert _rel(out, ref) < 1e-2, _rel(out, ref)


@pytest.mark.parametrize("N,d,heads,kind", [(4096, 40, 8, "ramp80"), (4096, 40, 8, "step"), (1035, 40, 8, "step"),
                                           (257, 64, 16, "step"), (1024, 80, 8, "ramp150"), (1024, 80, 8, "step"),
                                           (576, 160, 8, "step"), (4096, 40, 8, "late_step")])
def test_self_attention_overflow_is_recomputed_exactly(lib, N, d, heads, kind):
    """Logits that outrun the single-pass tiles' stale reference by more than its headroom (> 132 nats above every earlier
    key): round 1 documented the row as NaN.  Now the kernel notices (the row sum leaves its safe range), flags the work item,
    and the exact (running-maximum) variant launched right behind recomputes exactly those items: finite and equal to the fp32
    softmax (attention.py:217-229) -- which at such logits is a one-hot on the last keys.  B = 2 with only sample 0
    pathological: the re-run must touch flagged items only and leave the rest bit-identical to a plain run."""
    dev = torch.device("cuda:0")
    B = 2
    g = torch.Generator().manual_seed(11 + N + d)
    C = heads * d
    q = torch.randn(B, N, heads, d, generator=g)
    k = torch.randn(B, N, heads, d, generator=g) * 0.5
    u = torch.nn.functional.normalize(torch.randn(heads, d, generator=g), dim=-1)
    idx = torch.arange(N, dtype=torch.float32)
    if kind.startswith("ramp"):            # a sustained climb of 80 / 150 nats per 128 keys
        prof = idx / 128.0 * float(kind[4:])
    elif kind == "step":                    # flat, then +400 nats from the middle on
        prof = (idx >= N // 2).float() * 400.0
    else:                                   # the jump sits in the very last keys (the masked / final tile)
        prof = (idx >= N - 3).float() * 300.0
    comp = prof * math.sqrt(d) / 4.0
    q0 = q[0] - (q[0] * u).sum(-1, keepdim=True) * u + 4.0 * u
    k0 = k[0] - (k[0] * u).sum(-1, keepdim=True) * u + comp.view(N, 1, 1) * u
    q = torch.stack((q0, q[1]))
    k = torch.stack((k0, k[1]))
    q = q.reshape(B, N, C).to(dev).bfloat16()
    k = k.reshape(B, N, C).to(dev).bfloat16()
    v = torch.randn(B, N, C, generator=g).to(dev).bfloat16()
    qk = torch.cat((q, k), dim=-1).contiguous()
    Np = (N + 7) // 8 * 8
    vt = torch.zeros(B, C, Np, device=dev, dtype=torch.bfloat16)
    vt[:, :, :N] = v.transpose(1, 2)
    out = torch.zeros(B, N, C, device=dev, dtype=OP16())
    rc = lib.pbe_op_self_attention(qk.data_ptr(), vt.data_ptr(), out.data_ptr(), B, N, heads, d, _stream())
    assert rc == 0, _err(lib)
    torch.cuda.synchronize()

    def split(t):
        return t.float().view(B, N, heads, d).permute(0, 2, 1, 3)

    sim = torch.einsum("bhid,bhjd->bhij", split(q).double(), split(k).double()) * d ** -0.5
    ref = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), split(v).double()).permute(0, 2, 1, 3).reshape(B, N, C).float()
    assert torch.isfinite(out.float()).all()
    assert _rel(out[0], ref[0]) < 1e-2, _rel(out[0], ref[0])
    assert _rel(out[1], ref[1]) < 1e-2, _rel(out[1], ref[1])
    # the benign sample alone gives the same bits: the re-run did not touch it, and the flags were left clean
    out1 = torch.zeros(1, N, C, device=dev, dtype=OP16())
    rc = lib.pbe_op_self_attention(qk[1:].contiguous().data_ptr(), vt[1:].contiguous().data_ptr(), out1.data_ptr(), 1, N, heads, d,
                                   _stream())
    assert rc == 0, _err(lib)
    torch.cuda.synchronize()
    assert torch.equal(out1[0], out[1])


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("Nb,HW,C0,C1,eps,silu", [(2, 4096, 320, 0, 1e-5, 1), (2, 1024, 640, 320, 1e-5, 1),
                                                  (2, 256, 1280, 640, 1e-5, 1), (3, 64, 1280, 1280, 1e-5, 1),
                                                  (2, 1024, 640, 0, 1e-6, 0), (2, 16, 64, 128, 1e-5, 1)])
def test_groupnorm_concat_silu(lib, Nb, HW, C0, C1, eps, silu):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5 + C0 + C1)
    x0 = (torch.randn(Nb, HW, C0, generator=g) * 1.5 + 0.3).to(dev)
    x1 = (torch.randn(Nb, HW, C1, generator=g) * 0.7 - 0.2).to(dev) if C1 else None
    C = C0 + C1
    gamma = (1 + 0.1 * torch.randn(C, generator=g)).to(dev)
    beta = (0.1 * torch.randn(C, generator=g)).to(dev)
    y = torch.zeros(Nb, HW, C, device=dev, dtype=OP16())
    raw = torch.zeros(Nb, HW, C, device=dev, dtype=OP16())
    ws = torch.zeros(lib.pbe_op_groupnorm_workspace_bytes(Nb, HW) // 4 + 16, device=dev)
    rc = lib.pbe_op_groupnorm(x0.data_ptr(), C0, _p(x1), C1, Nb, HW, gamma.data_ptr(), beta.data_ptr(),
                              ctypes.c_float(eps), silu, y.data_ptr(), raw.data_ptr(), ws.data_ptr(), _stream())
    assert rc == 0, _err(lib)
    torch.cuda.synchronize()
    xc = torch.cat((x0, x1), dim=-1) if C1 else x0
    ref = F.group_norm(xc.permute(0, 2, 1), 32, gamma, beta, eps=eps).permute(0, 2, 1)
    if silu:
        ref = F.silu(ref)
    assert (y.float() - ref).abs().max().item() < 0.04 and _rel(y, ref) < 4e-3
    assert torch.equal(raw, xc.to(OP16()))


@pytest.mark.parametrize("M,C", [(8192, 320), (2048, 640), (515, 1280), (64, 64)])
def test_layernorm(lib, M, C):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(C)
    x = (torch.randn(M, C, generator=g) * 2 + 0.5).to(dev)
    gamma = (1 + 0.1 * torch.randn(C, generator=g)).to(dev)
    beta = (0.1 * torch.randn(C, generator=g)).to(dev)
    y = torch.zeros(M, C, device=dev, dtype=OP16())
    rc = lib.pbe_op_layernorm(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), M, C,
                              ctypes.c_float(1e-5), _stream())
    assert rc == 0, _err(lib)
    torch.cuda.synchronize()
    ref = F.layer_norm(x, (C,), gamma, beta, eps=1e-5)
    assert _rel(y, ref) < 4e-3


def test_upsample2x(lib):
    dev = torch.device("cuda:0")
    x = torch.randn(2, 8, 8, 128, device=dev)
    y = torch.zeros(2, 16, 16, 128, device=dev, dtype=OP16())
    rc = lib.pbe_op_upsample2x(x.data_ptr(), y.data_ptr(), 2, 8, 8, 128, _stream())
    assert rc == 0, _err(lib)
    torch.cuda.synchronize()
    ref = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1).to(OP16())
    assert torch.equal(y, ref)


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("order", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("cfg", [0, 1])
def test_sampler_step_bit_exact_vs_oracle(lib, order, cfg):
    """The fused K11 kernel equals the oracle restatement of plms.py:185-246 bit for bit (fp32, no FMA contraction)."""
    from oracle import sampler_ref as S
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(50 + order)
    B, shape = 3, (3, 4, 64, 64)
    eu, ec, h1, h2, h3, x = (torch.randn(shape, generator=g) for _ in range(6))
    tab = S.ddim_tables(S.make_schedule_buffers()["alphas_cumprod"], 50)
    index = 37
    a_t, a_prev, sig, s1m = (float(tab[k][index]) for k in ("alphas", "alphas_prev", "sigmas", "sqrt_one_minus_alphas"))
    scale = 5.0
    e = eu + scale * (ec - eu) if cfg else eu
    if order == 0:
        ep = e
    elif order == 1:
        ep = (3 * e - h1) / 2
    elif order == 2:
        ep = (23 * e - 16 * h1 + 5 * h2) / 12
    elif order == 3:
        ep = (55 * e - 59 * h1 + 37 * h2 - 9 * h3) / 24
    else:
        ep = (h1 + e) / 2
    xp_ref, x0_ref = S._x_prev_and_pred_x0(x, ep, a_t, a_prev, sig, s1m)
    d = lambda t: t.to(dev).contiguous()
    eu_d, ec_d, h1_d, h2_d, h3_d, x_d = map(d, (eu, ec, h1, h2, h3, x))
    e_out, xp, x0 = (torch.empty(shape, device=dev) for _ in range(3))
    rc = lib.pbe_sampler_step(eu_d.data_ptr(), ec_d.data_ptr() if cfg else None, ctypes.c_float(scale), cfg, order,
                              h1_d.data_ptr(), h2_d.data_ptr(), h3_d.data_ptr(), x_d.data_ptr(), ctypes.c_float(a_t),
                              ctypes.c_float(a_prev), ctypes.c_float(sig), ctypes.c_float(s1m), None, ctypes.c_float(1.0),
                              e_out.data_ptr(), xp.data_ptr(), x0.data_ptr(), x.numel(), _stream())
    assert rc == 0, _err(lib)
    torch.cuda.synchronize()
    assert torch.equal(e_out.cpu(), e)
    assert torch.equal(x0.cpu(), x0_ref)
    assert torch.equal(xp.cpu(), xp_ref)


def test_build_unet_input(lib):
    dev = torch.device("cuda:0")
    x, z, m = torch.randn(3, 4, 32, 32, device=dev), torch.randn(3, 4, 32, 32, device=dev), torch.rand(3, 1, 32, 32, device=dev)
    out = torch.empty(6, 9, 32, 32, device=dev)
    rc = lib.pbe_build_unet_input(x.data_ptr(), z.data_ptr(), m.data_ptr(), out.data_ptr(), 3, 4, 4, 1, 32 * 32, 2, _stream())
    assert rc == 0, _err(lib)
    torch.cuda.synchronize()
    ref = torch.cat([torch.cat((x, z, m), 1)] * 2)
    assert torch.equal(out, ref)


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,K,O,pre,act,res", [(16, 1280, 1280, 0, 0, False), (16, 320, 1280, 0, 1, False), (2, 768, 320, 0, 0, False),
                                               (32, 1280, 1280, 1, 0, True), (1, 1024, 4096, 0, 2, False), (8, 4096, 1024, 0, 0, True),
                                               (3, 100, 7, 1, 1, True), (17, 132, 33, 0, 2, False)])
def test_small_linear_matches_fp32_linear(lib, B, K, O, pre, act, res):
    """The fp32 GEMV of time_embed (openaimodel.py:623-628), the folded cross-attention (attention.py:207-230) and the
    one-token mapper (encoders/xf.py): K split across the warps of a CTA, fixed summation order."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(B * 7 + K + O)
    x = torch.randn(B, K, generator=g).to(dev)
    W = (torch.randn(O, K, generator=g) / math.sqrt(K)).to(dev)
    bias = (0.1 * torch.randn(O, generator=g)).to(dev)
    r = torch.randn(B, O, generator=g).to(dev) if res else None
    y = torch.full((B, O), float("nan"), device=dev)
    ys = torch.full((B, O), float("nan"), device=dev)
    rc = lib.pbe_op_small_linear(x.data_ptr(), W.data_ptr(), bias.data_ptr(), y.data_ptr(), B, K, O, pre, act, _p(r), ys.data_ptr(),
                                 _stream())
    assert rc == 0, _err(lib)
    torch.cuda.synchronize()
    ref = F.linear((F.silu(x) if pre else x).double(), W.double(), bias.double())
    ref = F.silu(ref) if act == 1 else (F.gelu(ref) if act == 2 else ref)
    if res:
        ref = ref + r.double()
    assert (y.double() - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
    assert (ys.double() - F.silu(ref)).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
    y2 = torch.empty_like(y)
    lib.pbe_op_small_linear(x.data_ptr(), W.data_ptr(), bias.data_ptr(), y2.data_ptr(), B, K, O, pre, act, _p(r), None, _stream())
    torch.cuda.synchronize()
    assert torch.equal(y, y2)          # fixed summation order: bit-identical from run to run
