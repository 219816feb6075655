#!/usr/bin/env python
"""Benchmark of the Paint-by-Example denoising hot path (BASELINE.json metric: 512² 50-step PLMS CFG images/sec).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

Workload (BASELINE.json configs[1]): v1.yaml U-Net (859.5 M parameters, seeded synthetic weights), batch 8 edit
requests per GPU (CFG doubles to 16), 64x64 latent (512x512 image), PLMS 50 steps (51 U-Net calls), guidance scale 5.
One "step" = one full `PLMSSampler.sample()` over one batch of 8 synthetic requests.  Multi-GPU = independent batches
per rank (weak scaling, no collective on the data path).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# NCCL's own log (communicator, rank count, transport) goes to stderr at INFO level so a driver can check the ranks;
# PBE_NCCL_DEBUG overrides.  stdout carries exactly ONE line, the JSON result: libraries write to file descriptor 1 behind
# Python's back (NCCL prints its log there), so fd 1 is pointed at stderr for the whole run and the JSON line goes to a
# private duplicate of the original stdout.
# (an NCCL_DEBUG preset by the image -- the GPU boxes export a quieter level -- is overridden: the rank check needs INFO)
os.environ["NCCL_DEBUG"] = os.environ.get("PBE_NCCL_DEBUG", "INFO")
_JSON_OUT = os.fdopen(os.dup(1), "w")
sys.stdout.flush()
os.dup2(2, 1)


def emit(line: dict) -> None:
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


import torch

F64_MIN = 771.3e9   # algorithmic FLOPs per sample-eval at 64x64 with the dead cross-attention work elided
# FLOPs the CFG-pair plan does NOT execute, per sample-eval at 64x64: the layers in front of the first cross-attention
# run once per (uncond, cond) pair -- conv_in 0.212 + first ResBlock 15.099 + proj_in 0.839 + qkv 2.517 + self-attention
# 21.475 = 40.142 GFLOP per pair, i.e. half of that per sample-eval.  Only executed work is claimed below.
F64_PAIR_SHARED = 40.142e9 / 2
# ... nor by the sub-pixel form of "nearest-2x upsample + 3x3 conv" (4/9 of the literal FLOPs) at the two levels whose phase
# launches fill the GPU at CFG batch >= 8: 16->32 (1280 ch) and 32->64 (640 ch), 30.2 GFLOP each in the literal form
F64_SUBPIXEL_SAVED = 2 * 30.199e9 * 5.0 / 9.0


def _ncu_traffic():
    """DRAM read + write bytes of the dominant kernel's representative launch, from the `ncu --set full` capture of the
    CURRENT build (tools/ncu_summary.py writes profiles/ncu_traffic.json next to the summary); None when no capture exists."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def unet_calls(args) -> int:
    """U-Net evaluations per image: PLMS evaluates the first step twice (plms.py:230-235), DDIM once per step."""
    return args.sampler_steps + 1 if args.sampler == "plms" else args.sampler_steps


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}",
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def _reference_root():
    """Where an importable copy of the unmodified reference lives on this box, if anywhere."""
    for cand in (os.environ.get("PBE_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "ldm", "modules", "diffusionmodules")):
            return cand
    return None


def cpu_reference_call_seconds(n_calls=1, threads=None, hw=64):
    """Time the reference's CPU implementation of one CFG U-Net call of ONE image (batch 2): the reference's own
    `UNetModel` when a copy of the reference is importable on this box (kind "reference"), otherwise the fp32 oracle port
    (kind "port"; torch.equal to the live reference in tests/test_oracle_pinned.py).  Returns (median s per call, threads, kind)."""
    from oracle import sampler_ref as S, unet_ref as U
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cfg = U.V1_CFG
    sd = U.make_state_dict(cfg, 321)
    req = S.synthetic_request(1, hw, hw, seed=321)
    x9 = torch.cat((req["x_T"], req["z_inpaint"], req["mask"]), 1)
    x_in = torch.cat([x9] * 2)
    c_in = torch.cat((req["uc"], req["c"]))
    t = torch.full((2,), 981, dtype=torch.int64)
    kind, fn = "port", (lambda: U.unet_forward(sd, cfg, x_in, t, c_in))
    root = _reference_root()
    if root is not None:
        try:
            os.environ["PBE_REFERENCE"] = root
            from oracle import reference_bridge as R
            ref = R.build_reference_unet(cfg, sd)

            def fn():
                with torch.no_grad():
                    return ref(x_in, t, context=c_in)
            kind = "reference"
        except Exception as e:   # an incomplete copy: say so on stderr and time the port
            print(f"[bench] reference at {root} not importable ({type(e).__name__}: {e}); timing the oracle port", file=sys.stderr)
    times = []
    for _ in range(n_calls):
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
    return statistics.median(times), threads, kind


def workload_string(args) -> str:
    """config.workload, shared by both arms (BASELINE configs[1] with the default flags)."""
    B, hw, Sn = args.batch, args.latent, args.sampler_steps
    calls = Sn + 1 if args.sampler == "plms" else Sn
    name = "BASELINE configs[1]: " if (B, hw, Sn, args.sampler) == (8, 64, 50, "plms") else ""
    return (f"{name}v1.yaml U-Net (859.5M params, seeded random weights), batch {B} per GPU (CFG batch {2 * B}), "
            f"{hw}x{hw} latent ({hw * 8}x{hw * 8} image), {args.sampler.upper()} {Sn} steps ({calls} U-Net calls), "
            f"guidance scale 5")


def base_config(args, world) -> dict:
    """`config` of the JSON line -- the same dict for every arm, so the driver compares like with like."""
    return {"workload": workload_string(args), "global_batch": world * args.batch,
            "parallelism": f"dp{world} (independent requests, no collective)",
            "l2": "working set per step (1.7 GB bf16 weights + >1 GB activations per U-Net call) exceeds the "
                  "126 MB L2; no explicit flush"}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores (all threads).  Each "step"
    is a bounded sample of the workload: ONE CFG U-Net call of one image (batch 2) at the workload's latent size; images/s
    = 1 / (calls per image x median call time), the way BASELINE.md section 4 extrapolates.  Rank 0 only under torchrun."""
    if rank != 0:
        return
    total = args.steps + args.warmup
    calls = unet_calls(args)
    per_call, threads, kind = cpu_reference_call_seconds(n_calls=max(1, total), hw=args.latent)
    ips = 1.0 / (calls * per_call)
    what = ("the reference's own ldm UNetModel.forward (fp32, CPU)" if kind == "reference"
            else "the fp32 oracle port of ldm UNetModel.forward (CPU)")
    line = {
        "impl": "reference", "metric": "images_per_sec_512px_plms50_cfg", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_call * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config(args, args.gpus),
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": kind,
                         "sample": f"median of {max(1, total)} CFG U-Net call(s) of ONE image (batch 2, {args.latent}x{args.latent} "
                                   f"latent) through {what} on {threads} host threads x {calls} calls per image "
                                   f"(extrapolated); {per_call:.2f} s per call; ms_per_step is that call time",
                         "arm_batch": 2, "unet_calls_per_image": calls},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_torch_eager(args, rank, world):
    """--impl torch-eager: the GPU-side bar the reference itself would set on this box (SURVEY.md
    section 2.2): the reference algorithm (fp32 oracle restatement of UNetModel.forward) through PyTorch library kernels on
    the same B200 -- eager fp32 (TF32 off / on), under torch.autocast as scripts/inference.py:301-303 runs it (fp16) and with
    bf16, and with F.scaled_dot_product_attention in place of the literal softmax(QK^T)V -- per U-Net call at CFG batch 2 and
    at the workload's CFG batch.  Not the product and not the reference arm: a context line for profiles/."""
    if rank != 0:
        return
    import torch.nn.functional as F
    from oracle import unet_ref as U
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    cfg = U.V1_CFG
    sd = {k: v.to(dev) for k, v in U.make_state_dict(cfg, 321).items()}
    hw = args.latent
    literal_attention = U.cross_attention

    def sdpa_attention(sd_, p, x, context, heads):
        q = F.linear(x, sd_[p + ".to_q.weight"])
        ctx = x if context is None else context
        k, v = F.linear(ctx, sd_[p + ".to_k.weight"]), F.linear(ctx, sd_[p + ".to_v.weight"])
        b, n, c = q.shape
        sp = lambda t_: t_.view(b, t_.shape[1], heads, c // heads).transpose(1, 2)
        o = F.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(b, n, c)
        return F.linear(o, sd_[p + ".to_out.0.weight"], sd_[p + ".to_out.0.bias"])

    def time_variant(Bc, tf32, autocast_dtype, sdpa):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        U.cross_attention = sdpa_attention if sdpa else literal_attention
        g = torch.Generator().manual_seed(Bc)
        x = torch.randn(Bc, 9, hw, hw, generator=g).to(dev)
        t = torch.full((Bc,), 501, dtype=torch.int64, device=dev)
        c = torch.randn(Bc, 1, 768, generator=g).to(dev)

        def call():
            if autocast_dtype is None:
                return U.unet_forward(sd, cfg, x, t, c)
            with torch.autocast("cuda", dtype=autocast_dtype):
                return U.unet_forward(sd, cfg, x, t, c)
        try:
            for _ in range(2):
                call()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = max(3, args.steps)
            e0.record()
            for _ in range(n):
                call()
            e1.record()
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1) / n
        except torch.cuda.OutOfMemoryError:
            torch.cuda.empty_cache()
            return None
        finally:
            U.cross_attention = literal_attention

    rows = []
    for Bc in (2, 2 * args.batch):
        for name, tf32, ac, sdpa in (("fp32 eager, TF32 off (the oracle as the parity tests run it)", False, None, False),
                                     ("fp32 eager, TF32 on (PyTorch defaults for conv)", True, None, False),
                                     ("fp32 eager + SDPA, TF32 on", True, None, True),
                                     ("autocast fp16 (scripts/inference.py:301-303)", True, torch.float16, False),
                                     ("autocast fp16 + SDPA", True, torch.float16, True),
                                     ("autocast bf16 + SDPA", True, torch.bfloat16, True)):
            ms = time_variant(Bc, tf32, ac, sdpa)
            rows.append({"cfg_batch": Bc, "variant": name, "ms_per_unet_call": ms})
            print(f"[torch-eager] CFG batch {Bc:3d} {name}: {ms if ms is None else round(ms, 2)} ms", file=sys.stderr)
    calls = unet_calls(args)
    best = min((r["ms_per_unet_call"] for r in rows if r["cfg_batch"] == 2 * args.batch and r["ms_per_unet_call"]), default=None)
    emit({"impl": "torch-eager", "metric": "images_per_sec_512px_plms50_cfg",
          "value": (args.batch / (calls * best * 1e-3)) if best else None, "unit": "images/s", "n_gpus": 1,
          "higher_is_better": True, "dtype": "fp32 / fp16 / bf16 (see variants)", "data": "synthetic",
          "config": base_config(args, 1),
          "note": "value = batch / (U-Net calls per image x the FASTEST variant's call time at the workload's CFG batch): U-Net "
                  "calls only, no sampler arithmetic -- an upper bound for a PyTorch-library implementation of the reference",
          "variants": rows})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-eager"])
    ap.add_argument("--batch", type=int, default=8, help="edit requests per GPU per step (CFG doubles it)")
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--sampler-steps", type=int, default=50)
    ap.add_argument("--sampler", default="plms", choices=["plms", "ddim"],
                    help="plms (BASELINE configs 1-3, 5) or ddim (config 4: DDIM-20 latency)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default=None, help="write the per-op breakdown of one U-Net call to this JSON")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.impl == "torch-eager":
        return run_torch_eager(args, rank, world)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        # the synthetic weights are generated on the host by every rank: share the cores instead of oversubscribing them
        torch.set_num_threads(max(1, (os.cpu_count() or world) // world))

    from oracle import sampler_ref as S, unet_ref as U   # synthetic weights / requests only (not timed, not the product)
    from pbe_b200.diffusion import LatentDiffusion
    from pbe_b200.samplers import DDIMSampler, PLMSSampler

    cfg = U.V1_CFG
    sd = U.make_state_dict(cfg, 321)
    model = LatentDiffusion(unet_config=dict(params=dict(cfg)))
    model.load_state_dict({"model.diffusion_model." + k: v for k, v in sd.items()}, strict=False)
    model = model.to(dev).eval()
    del sd
    B, hw, Sn = args.batch, args.latent, args.sampler_steps
    req = S.synthetic_request(B, hw, hw, seed=321 + rank)
    pin = {k: v.contiguous().pin_memory() for k, v in req.items()}
    d = {k: v.to(dev) for k, v in req.items()}
    sampler = PLMSSampler(model) if args.sampler == "plms" else DDIMSampler(model)
    calls = unet_calls(args)
    flops_per_eval = F64_MIN * (hw * hw) / 4096.0 if hw == 64 else None   # SURVEY 8(d) gives F only for 64 and 96
    if hw == 96:
        flops_per_eval = 2079.9e9
    unet = model.model.diffusion_model
    from pbe_b200 import _lib as _pl
    fmt_f16 = bool(_pl.load().pbe_get_operand_format())

    def sample_device():
        out, _ = sampler.sample(S=Sn, conditioning=d["c"], batch_size=B, shape=[4, hw, hw], verbose=False,
                                unconditional_guidance_scale=5.0, unconditional_conditioning=d["uc"], eta=0.0,
                                x_T=d["x_T"], test_model_kwargs=dict(images_inpaint=d["z_inpaint"], images_mask=d["mask"]))
        return out

    host_out = torch.empty((B, 4, hw, hw), dtype=torch.float32).pin_memory()

    def sample_e2e():
        dd = {k: v.to(dev, non_blocking=True) for k, v in pin.items()}
        out, _ = sampler.sample(S=Sn, conditioning=dd["c"], batch_size=B, shape=[4, hw, hw], verbose=False,
                                unconditional_guidance_scale=5.0, unconditional_conditioning=dd["uc"], eta=0.0,
                                x_T=dd["x_T"], test_model_kwargs=dict(images_inpaint=dd["z_inpaint"], images_mask=dd["mask"]))
        host_out.copy_(out, non_blocking=True)
        return out

    h2d = sum(v.numel() * v.element_size() for v in pin.values())
    d2h = host_out.numel() * 4

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(args.warmup, 3)):
        sample_device()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    total_ms = timed(sample_device, args.steps)
    for _ in range(1):
        sample_e2e()
    e2e_ms = timed(sample_e2e, args.steps)
    clk = clocks.stop() if rank == 0 else None

    # ---- U-Net step latency (p50 / p99) and per-kernel-family breakdown of one call ----
    Bc = 2 * B
    x_in = torch.randn(Bc, 9, hw, hw, device=dev)
    t_in = torch.full((Bc,), 501, device=dev, dtype=torch.int64)
    unet.set_context(torch.cat((d["uc"].expand(B, 1, 768), d["c"])))
    eps = torch.empty(Bc, 4, hw, hw, device=dev)
    lat = []
    for i in range(60):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        unet.run(x_in, t_in, out=eps)
        e1.record()
        torch.cuda.synchronize(dev)
        if i >= 10:
            lat.append(e0.elapsed_time(e1))
    lat.sort()
    prof = unet.profile(x_in, t_in)
    prof = unet.profile(x_in, t_in)   # second pass: warm
    fam = {}
    for r in prof:
        f = fam.setdefault(r["family"], dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
        f["ms"] += r["ms"]; f["flops"] += r["flops"]; f["bytes"] += r["bytes"]; f["launches"] += 1
    prof_total = sum(r["ms"] for r in prof)
    peaks = _peaks()
    if rank == 0 and args.profile_out:
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
        json.dump(dict(ops=prof, families=fam, total_ms=prof_total, Bc=Bc, hw=hw), open(args.profile_out, "w"), indent=1)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    imgs = world * B * args.steps
    value = imgs / (total_ms / 1e3)
    e2e_value = imgs / (e2e_ms / 1e3)
    launches = calls * (unet.launches_per_forward() + 2) + 1
    gemm = fam.get("conv_gemm", dict(ms=1.0, flops=0.0, launches=1))
    gemm_tf = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:   # the CPU baseline is reported at N = 1 only
        per_call, threads, kind = cpu_reference_call_seconds(n_calls=5, hw=hw)
        cpu_baseline = {"value": 1.0 / (calls * per_call), "unit": "images/s", "cores": threads, "kind": kind,
                        "sample": f"median of 5 CFG U-Net calls of ONE image (batch 2, {hw}x{hw} latent) of the "
                                  f"{'reference UNetModel' if kind == 'reference' else 'fp32 oracle port'} x {calls} calls per "
                                  f"image (extrapolated); {per_call:.2f} s per call",
                        "arm_batch": 2, "unet_calls_per_image": calls}
    traffic = _ncu_traffic()
    line = {
        "metric": "images_per_sec_512px_plms50_cfg", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp16" if fmt_f16 else "bf16", "data": "synthetic",
        "dtype_note": ("fp16 tensor-core operands (saturating conversions) and 16-bit fp16 residual stream, fp32 accumulation and "
                       "statistics; bf16 Q/K/V/P inside self-attention -- the precision the reference runs at under torch.autocast"
                       if fmt_f16 else "bf16 tensor-core operands, fp32 accumulation, statistics and residual stream (PBE_OPERANDS=bf16)"),
        "config": base_config(args, world),
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches * args.steps,
        "unet_step_ms": {"p50": lat[len(lat) // 2], "p99": lat[min(len(lat) - 1, int(len(lat) * 0.99))],
                         "cfg_batch": Bc, "floor_ms_at_sustained_peak": (Bc * flops_per_eval / (peaks["tf_sustained"] * 1e12) * 1e3) if flops_per_eval else None},
        "tensor_utilisation_whole_job": {"achieved_tflops": (B * args.steps * 2 * calls * (flops_per_eval - ((F64_PAIR_SHARED + (F64_SUBPIXEL_SAVED if (fmt_f16 and B >= 8) else 0.0)) if hw == 64 else 0.0))
                                                             / (total_ms / 1e3) / 1e12) if flops_per_eval else None,
                                         "note": "executed FLOPs only: the CFG-pair plan evaluates the layers in front of the "
                                                 "first cross-attention once per (uncond, cond) pair, and two of the three "
                                                 "upsample convs run in their sub-pixel form (4/9 of the literal FLOPs)",
                                         "peak_tflops_sustained": peaks["tf_sustained"], "peak_source": peaks["src"]},
        "roofline": {"bound": "tensor", "kernel": "conv_gemm_kernel (implicit-GEMM conv / linear, tcgen05)",
                     "achieved": gemm_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                     "frac": gemm_tf / peaks["tf_sustained"],
                     "traffic": (traffic or {}).get("dram_bytes"),
                     "traffic_note": (traffic or {}).get("note", "no ncu --set full capture of the current build under profiles/"),
                     "how": f"sum of algorithmic FLOPs of the {gemm['launches']} conv_gemm launches of one U-Net call "
                            f"(CFG batch {Bc}) / sum of their CUDA-event durations (eager pass after the timed region); "
                            f"peak = {peaks['src']} sustained bf16",
                     "share_of_unet_call": gemm["ms"] / prof_total},
        "kernel_families_ms_per_unet_call": {k: round(v["ms"], 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])},
        "cpu_baseline": cpu_baseline,
        "clocks": clk,
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
