"""CPU-only checks (no GPU): the C-ABI library loads and exports every declared symbol, the drop-in module exposes the
reference's state-dict layout, the samplers build the reference's schedule tables, key spellings / error conventions
match the reference, and the product refuses to run without CUDA (no fallback)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import sampler_ref as S
from oracle import unet_ref as U


def test_library_loads_and_exports_declared_symbols():
    from pbe_b200 import _lib
    lib = _lib.load()
    syms = _lib.exported_symbols()
    assert len(syms) >= 19
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.pbe_last_error() is not None


def test_unet_state_dict_keys_match_reference(golden_dir):
    from pbe_b200.unet import UNetModel
    m = UNetModel(**U.SMALL_CFG)
    idx = json.load(open(os.path.join(golden_dir, "golden_index.json")))
    assert sorted(m.state_dict().keys()) == idx["state_dict_keys"]["keys"]
    shapes = U.param_shapes(U.SMALL_CFG)
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(shapes[k]), k
    big = UNetModel.__new__(UNetModel)  # shapes only, no 3.4 GB allocation
    from pbe_b200.unet import unet_param_shapes
    s = unet_param_shapes(**{k: U.V1_CFG[k] for k in ("in_channels", "out_channels", "model_channels", "num_res_blocks",
                                                      "channel_mult", "attention_resolutions", "num_heads",
                                                      "context_dim")})
    assert len(s) == 686 and sum(int(np.prod(v)) for v in s.values()) == 859_535_364


def test_checkpoint_style_loading_under_model_diffusion_model_prefix():
    from pbe_b200.diffusion import LatentDiffusion
    sd = U.make_state_dict(U.SMALL_CFG, 1)
    m = LatentDiffusion(unet_config=dict(target="ldm.modules.diffusionmodules.openaimodel.UNetModel",
                                         params=dict(U.SMALL_CFG)))
    r = m.load_state_dict({"model.diffusion_model." + k: v for k, v in sd.items()}, strict=False)
    assert not r.unexpected_keys
    assert not [k for k in r.missing_keys if k.startswith("model.")]
    got = m.model.diffusion_model.state_dict()
    assert torch.equal(got["input_blocks.0.0.weight"], sd["input_blocks.0.0.weight"])
    assert m.num_timesteps == 1000 and m.alphas_cumprod.dtype == torch.float32


def test_no_cpu_fallback():
    from pbe_b200.diffusion import LatentDiffusion
    from pbe_b200.samplers import PLMSSampler
    m = LatentDiffusion(unet_config=dict(params=dict(U.SMALL_CFG)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.apply_model(torch.zeros(1, 9, 32, 32), torch.zeros(1, dtype=torch.int64), torch.zeros(1, 1, 768))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PLMSSampler(m).sample(S=4, batch_size=1, shape=[4, 32, 32], conditioning=torch.zeros(1, 1, 768), verbose=False,
                              x_T=torch.zeros(1, 4, 32, 32),
                              test_model_kwargs=dict(images_inpaint=torch.zeros(1, 4, 32, 32),
                                                     images_mask=torch.zeros(1, 1, 32, 32)))


def test_unsupported_unet_options_are_rejected_loudly():
    from pbe_b200.unet import UNetModel
    with pytest.raises(NotImplementedError):
        UNetModel(**dict(U.SMALL_CFG, use_scale_shift_norm=True))
    with pytest.raises(NotImplementedError):
        UNetModel(**dict(U.SMALL_CFG, transformer_depth=2))


@pytest.mark.parametrize("S_steps", [50, 20, 8])
def test_sampler_schedule_tables_match_oracle(S_steps):
    from pbe_b200.diffusion import LatentDiffusion
    from pbe_b200.samplers import DDIMSampler, PLMSSampler
    m = LatentDiffusion(unet_config=dict(params=dict(U.SMALL_CFG)))
    buf = S.make_schedule_buffers()
    assert torch.equal(m.alphas_cumprod, buf["alphas_cumprod"]) and torch.equal(m.betas, buf["betas"])
    tab = S.ddim_tables(buf["alphas_cumprod"], S_steps)
    for cls in (PLMSSampler, DDIMSampler):
        smp = cls(m)
        smp.make_schedule(S_steps, ddim_eta=0.0, verbose=False)
        assert list(smp.ddim_timesteps) == list(tab["timesteps"])
        assert smp._coef["a_t"] == [float(v) for v in tab["alphas"]]
        assert smp._coef["a_prev"] == [float(v) for v in tab["alphas_prev"]]
        assert smp._coef["sqrt_one_minus_at"] == [float(v) for v in tab["sqrt_one_minus_alphas"]]
        assert all(v == 0.0 for v in smp._coef["sigma"])
        assert isinstance(smp.ddim_alphas_prev, np.ndarray)        # reference keeps this one as numpy (util.py:66)
        assert smp.ddim_alphas.dtype == torch.float32


def test_plms_rejects_eta_and_ddim_rejects_batch_mismatch():
    from pbe_b200.diffusion import LatentDiffusion
    from pbe_b200.samplers import DDIMSampler, PLMSSampler
    m = LatentDiffusion(unet_config=dict(params=dict(U.SMALL_CFG)))
    with pytest.raises(ValueError, match="ddim_eta must be 0 for PLMS"):      # plms.py:25-26
        PLMSSampler(m).make_schedule(10, ddim_eta=0.5, verbose=False)
    with pytest.raises(ValueError):                                            # ddim.py:99-106
        DDIMSampler(m).sample(S=4, batch_size=2, shape=[4, 32, 32], conditioning=torch.zeros(1, 1, 768), verbose=False)
    # ddim.py:262-283: decode() forwards no inpainting kwargs, so p_sample_ddim's check (:203-204) always fires in this fork
    with pytest.raises(Exception, match="kwargs must contain either 'test_model_kwargs' or 'rest' key"):
        DDIMSampler(m).decode(torch.zeros(1, 4, 32, 32), torch.zeros(1, 1, 768), 3)


def test_inpaint_kwargs_spellings():
    from pbe_b200.samplers import _inpaint_kwargs
    z, mk = torch.zeros(1, 4, 8, 8), torch.ones(1, 1, 8, 8)
    for kw in (dict(test_model_kwargs=dict(images_inpaint=z, images_mask=mk)),
               dict(test_model_kwargs=dict(inpaint_image=z, inpaint_mask=mk)),
               dict(rest=torch.cat((z, mk), 1))):
        a, b = _inpaint_kwargs(kw)
        assert a.shape == z.shape and b.shape == mk.shape
    assert _inpaint_kwargs({}) is None
    with pytest.raises(KeyError):
        _inpaint_kwargs(dict(test_model_kwargs={}))


def test_two_rank_request_sharding_gloo(tmp_path):
    """Multi-GPU path = independent request shards: world_size-2 gloo run of the sharding helper."""
    import torch.multiprocessing as mp
    from pbe_b200.sharding import shard_requests
    assert shard_requests(10, 0, 4) == [0, 4, 8] and shard_requests(10, 3, 4) == [3, 7]
    assert sorted(sum((shard_requests(3500, r, 8) for r in range(8)), [])) == list(range(3500))
    mp.spawn(_gloo_worker, args=(2, str(tmp_path)), nprocs=2, join=True)
    got = sorted(int(x) for r in range(2) for x in open(os.path.join(tmp_path, f"r{r}.txt")).read().split())
    assert got == list(range(7))


def _gloo_worker(rank, world, out_dir):
    import torch.distributed as dist
    from pbe_b200.sharding import gather_latents, shard_requests
    dist.init_process_group("gloo", init_method=f"file://{out_dir}/rdzv", rank=rank, world_size=world)
    mine = shard_requests(7, rank, world)
    lat = torch.stack([torch.full((4, 2, 2), float(i)) for i in mine]) if mine else torch.zeros(0, 4, 2, 2)
    ids, full = gather_latents(mine, lat, 7)
    assert ids == list(range(7))
    assert all(float(full[i, 0, 0, 0]) == float(i) for i in range(7))
    open(os.path.join(out_dir, f"r{rank}.txt"), "w").write(" ".join(str(i) for i in (mine if rank == 0 else mine)))
    dist.destroy_process_group()


def test_vae_state_dict_keys_and_no_cpu_fallback(golden_dir):
    """pbe_b200.AutoencoderKL carries the reference's encoder / decoder keys (a checkpoint's first_stage_model.* entries
    load with strict=False) and refuses to encode / decode without a CUDA device."""
    import json
    import os
    import pytest
    import torch
    from pbe_b200.vae import AutoencoderKL
    dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
              num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    m = AutoencoderKL(ddconfig=dd, embed_dim=4, lossconfig=dict(target="torch.nn.Identity"), monitor="val/rec_loss")
    idx = json.load(open(os.path.join(golden_dir, "golden_index.json")))
    assert sorted(m.state_dict().keys()) == idx["vae_state_dict_keys"]["keys"]
    sd = {"first_stage_model." + k: torch.zeros_like(v) for k, v in m.state_dict().items()}
    sd["first_stage_model.loss.logvar"] = torch.zeros(())     # training-only entries of a checkpoint: ignored
    host = torch.nn.Module()
    host.first_stage_model = m
    missing, unexpected = host.load_state_dict(sd, strict=False)
    assert not missing and unexpected == ["first_stage_model.loss.logvar"]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.decode(torch.zeros(1, 4, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.encode(torch.zeros(1, 3, 64, 64))
    with pytest.raises(NotImplementedError):
        AutoencoderKL(ddconfig=dict(dd, attn_resolutions=[16]), embed_dim=4)


def test_clip_embedder_state_dict_keys_and_no_cpu_fallback(golden_dir):
    import json
    import os
    import pytest
    import torch
    from pbe_b200.clip import FrozenCLIPImageEmbedder
    m = FrozenCLIPImageEmbedder()
    idx = json.load(open(os.path.join(golden_dir, "golden_index.json")))
    assert sorted(m.state_dict().keys()) == idx["clip_state_dict_keys"]["keys"]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 224, 224))
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 200, 224))


def test_bench_reference_arm_prints_one_json_line():
    """The measurement contract: `bench.py --impl reference` times the reference's CPU implementation of the path (its own
    UNetModel where /root/reference is mounted -- kind "reference" --, the oracle port otherwise -- kind "port", one of the
    places allowed to execute oracle/) on the host cores and prints exactly ONE line on stdout, a JSON object carrying the
    base keys, the arm's own cpu_baseline and an e2e block with zero copy bytes; everything else goes to stderr."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "images_per_sec_512px_plms50_cfg" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["config"]["workload"].startswith("BASELINE configs[1]") and "model" not in d["config"]
    have_ref = os.path.isdir(os.path.join(os.environ.get("PBE_REFERENCE", "/root/reference"), "ldm"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["unet_calls_per_image"] == 51 and d["cpu_baseline"]["arm_batch"] == 2
    assert set(d["config"]) == {"workload", "global_batch", "parallelism", "l2"}     # the same keys as our arm's config
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_preprocess_mirror_has_no_cpu_path_and_keeps_reference_errors():
    """pbe_b200.preprocess mirrors get_tensor / get_tensor_clip / Resize (scripts/inference.py:106-124,332): CPU tensors and
    wrong dtypes are refused loudly, torchvision's zero-std error is kept."""
    from pbe_b200 import preprocess as P
    with pytest.raises(RuntimeError, match="no CPU path"):
        P.get_tensor()(torch.zeros(8, 8, 3, dtype=torch.uint8))
    with pytest.raises(TypeError):
        P.get_tensor_clip()(torch.zeros(8, 8, 3))
    with pytest.raises(RuntimeError, match="no CPU path"):
        P.prepare_inpaint(torch.zeros(8, 8, 3, dtype=torch.uint8), torch.zeros(8, 8, dtype=torch.uint8))
    with pytest.raises(RuntimeError, match="no CPU path"):
        P.Resize([4, 4])(torch.zeros(1, 1, 8, 8))
    with pytest.raises(ValueError, match="std evaluated to zero"):
        P._Normalize((0.5, 0.5, 0.5), (0.5, 0.0, 0.5))
    with pytest.raises(NotImplementedError):
        P.Resize(64)


def test_cited_reference_lines_exist():
    """Every `file.py:line[-line]` citation in the C header, the kernels, the host mirror, the oracle and the design documents
    points at lines that exist in the reference tree (authoring container only; the tree is not on the GPU box)."""
    import glob
    import re
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("/root/reference not mounted")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    by_name = {}
    for path in glob.glob(os.path.join(ref, "**", "*.py"), recursive=True):
        by_name.setdefault(os.path.basename(path), []).append(path)
    sources = [os.path.join(root, "include", "pbe_b200.h")] + glob.glob(os.path.join(root, "pbe_b200", "*.py")) + \
        glob.glob(os.path.join(root, "oracle", "*.py")) + glob.glob(os.path.join(root, "pbe_b200", "csrc", "*.*")) + \
        [os.path.join(root, "DESIGN.md"), os.path.join(root, "INTEGRATION.md")]
    pat = re.compile(r"([A-Za-z_0-9/]+\.py):(\d+)(?:-(\d+))?")
    checked, bad = 0, []
    for src in sources:
        for m in pat.finditer(open(src).read()):
            name, lo, hi = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            cands = [p for p in by_name.get(os.path.basename(name), []) if p.endswith(name)]
            if not cands:
                continue        # not a reference file (e.g. a citation of this repository or of a third-party library)
            n_lines = max(sum(1 for _ in open(p)) for p in cands)
            checked += 1
            if not (1 <= lo <= hi <= n_lines):
                bad.append((os.path.relpath(src, root), m.group(0), n_lines))
    assert checked > 50 and not bad, bad


def test_committed_bench_lines_follow_the_contract():
    """profiles/ holds the evidence the design document quotes: the final bench lines must carry every key of the measurement
    contract (metric / value / e2e with copy bytes / roofline / cpu_baseline / clocks / gpu_launches), and the ncu launch
    list must summarise (tools/launch_summary.py)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prof = os.path.join(root, "profiles")
    d = json.load(open(os.path.join(prof, "r01_bench_v26.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["metric"] == "images_per_sec_512px_plms50_cfg" and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("BASELINE configs[1]") and "model" not in d["config"]
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] > 0
    assert d["gpu_launches"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    d2 = json.load(open(os.path.join(prof, "r01_bench_v26_2gpu.json")))
    assert d2["n_gpus"] == 2 and d2["cpu_baseline"] is None and 1.8 < d2["value"] / d["value"] < 2.2
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "launch_summary.py"), os.path.join(prof, "r01_launches_v23.csv")],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "conv_gemm_kernel" in out.stdout and "flash_attn2_kernel" in out.stdout


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: no module of the product package may import it (statically checked)."""
    import ast
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for path in glob.glob(os.path.join(root, "pbe_b200", "**", "*.py"), recursive=True):
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n == "oracle" or n.startswith("oracle.") for n in names), path


def test_layernorm_fold_algebra():
    """The identity behind the LayerNorm fold of DESIGN.md 3.3 (Engine::fold_layernorm, csrc/engine.cu), in fp64 on the host:
    with W''[o][i] = W[o][i] g[i] - mean_i(W[o][:] g) and b'[o] = b[o] + sum_i W[o][i] beta[i],
        Linear(LayerNorm(x)) = rstd * (W'' x_raw) + b'          (attention.py:270-276, norm1 -> to_q/k/v, norm3 -> ff.net.0.proj)
    for rows whose mean is far from zero too (the centred rows make the mean term vanish instead of cancelling it), and the
    treatment of beta's image under q | k | v: k's share is constant over the keys of a query row (softmax-invariant), v's
    share passes through the attention and becomes W_out b_v (attention.py:207-230)."""
    g = torch.Generator().manual_seed(3)
    C, O, M = 64, 48, 37
    x = torch.randn(M, C, generator=g, dtype=torch.float64) * 0.7 + torch.randn(M, 1, generator=g, dtype=torch.float64) * 25.0
    W = torch.randn(O, C, generator=g, dtype=torch.float64) / C ** 0.5
    b = torch.randn(O, generator=g, dtype=torch.float64)
    gamma = 1.0 + 0.3 * torch.randn(C, generator=g, dtype=torch.float64)
    beta = 0.2 * torch.randn(C, generator=g, dtype=torch.float64)
    ref = torch.nn.functional.linear(torch.nn.functional.layer_norm(x, (C,), gamma, beta, 1e-5), W, b)
    Wg = W * gamma
    W2 = Wg - Wg.mean(dim=1, keepdim=True)
    b2 = b + W @ beta
    assert W2.sum(dim=1).abs().max() < 1e-12
    s, q = x.sum(dim=1), (x * x).sum(dim=1)                       # the producing epilogue's row statistics
    mean = s / C
    rstd = torch.rsqrt(q / C - mean * mean + 1e-5)
    out = rstd[:, None] * (x @ W2.T) + b2
    assert (out - ref).abs().max() < 1e-9 * ref.abs().max() * 25.0 ** 2  # the variance formula loses mean^2 / var digits, nothing else

    # q | k | v: dropping b_k and moving b_v behind the attention leaves softmax(q k^T) v W_out^T unchanged
    d = 16
    Wq, Wk, Wv, Wo = (torch.randn(d, C, generator=g, dtype=torch.float64) / C ** 0.5 for _ in range(4))
    Wo = torch.randn(C, d, generator=g, dtype=torch.float64) / d ** 0.5
    ln = torch.nn.functional.layer_norm(x, (C,), gamma, beta, 1e-5)
    qf, kf, vf = ln @ Wq.T, ln @ Wk.T, ln @ Wv.T
    att_ref = torch.softmax(qf @ kf.T * d ** -0.5, dim=-1) @ vf @ Wo.T
    z = (x - mean[:, None]) * rstd[:, None] * gamma                # LayerNorm without beta
    bq, bv = Wq @ beta, Wv @ beta
    q2, k2, v2 = z @ Wq.T + bq, z @ Wk.T, z @ Wv.T                 # k without its bias, v without its bias
    att = torch.softmax(q2 @ k2.T * d ** -0.5, dim=-1) @ v2 @ Wo.T + Wo @ bv
    assert (att - att_ref).abs().max() < 1e-9


def test_ff_proj_out_merge_algebra():
    """The identity behind the merged ff.net.2 / proj_out GEMM of DESIGN.md 3.6 (Engine::finalize, csrc/engine.cu), in fp64 on the
    host: nothing non-linear sits between the feed-forward's output Linear, its residual add and the SpatialTransformer's
    proj_out (attention.py:276 `x = self.ff(self.norm3(x)) + x`, 335-336 `x = self.proj_out(x); return x + x_in`), so
        proj_out(ff2(g) + x2) + x_in = [ W_po W_ff2 | W_po ] [ g | x2 ]^T + (W_po b_ff2 + b_po) + x_in
    -- one GEMM with K = 4C + C over the buffer that holds the GEGLU output and x2 side by side."""
    g = torch.Generator().manual_seed(5)
    C, M = 32, 29
    gg = torch.randn(M, 4 * C, generator=g, dtype=torch.float64)
    x2 = torch.randn(M, C, generator=g, dtype=torch.float64)
    x_in = torch.randn(M, C, generator=g, dtype=torch.float64)
    Wf = torch.randn(C, 4 * C, generator=g, dtype=torch.float64) / (4 * C) ** 0.5
    bf = torch.randn(C, generator=g, dtype=torch.float64)
    Wp = torch.randn(C, C, generator=g, dtype=torch.float64) / C ** 0.5
    bp = torch.randn(C, generator=g, dtype=torch.float64)
    lin = torch.nn.functional.linear
    ref = lin(lin(gg, Wf, bf) + x2, Wp, bp) + x_in
    Wm = torch.cat([Wp @ Wf, Wp], dim=1)
    bm = Wp @ bf + bp
    out = lin(torch.cat([gg, x2], dim=1), Wm, bm) + x_in
    assert Wm.shape == (C, 5 * C)
    assert (out - ref).abs().max() < 1e-12 * max(1.0, ref.abs().max().item())
