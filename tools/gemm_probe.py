"""Run a few representative GEMM shapes of the U-Net (CFG batch 16) through pbe_op_conv_gemm; used under ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pbe_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream

def case(name, Nb, H, W, C, k, Cout, mode=0, residual=False, rowbias=False, f32=True, b16=False, bn=0, iters=3):
    x = torch.randn(Nb, H, W, C, device=dev).bfloat16()
    w = (torch.randn(k * k, Cout, C, device=dev) / (C * k * k) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device=dev)
    ncol = Cout // 2 if mode == 1 else Cout
    res = torch.randn(Nb, H, W, ncol, device=dev) if residual else None
    rb = torch.randn(Nb, Cout, device=dev) if rowbias else None
    of = torch.empty(Nb, H, W, ncol, device=dev) if f32 else None
    ob = torch.empty(Nb, H, W, ncol, device=dev, dtype=torch.bfloat16) if (b16 or mode == 1) else None
    p = _lib.ptr
    if os.environ.get("PBE_PROBE_STATS") and mode == 0 and f32:
        stats = torch.zeros(Nb * H * W // 32 * Cout * 2 + 64, device=dev)
        lib.pbe_debug_set_gemm_stats_out(p(stats))
        name += " +stats"
    else:
        lib.pbe_debug_set_gemm_stats_out(None)
    for _ in range(iters):
        rc = lib.pbe_op_conv_gemm(p(x), Nb, H, W, C, k, 1, p(w), Cout, mode, p(bias), p(rb), p(res), p(of), p(ob), None, 0, bn, st)
        assert rc == 0, lib.pbe_last_error()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.pbe_op_conv_gemm(p(x), Nb, H, W, C, k, 1, p(w), Cout, mode, p(bias), p(rb), p(res), p(of), p(ob), None, 0, bn, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * Nb * H * W * Cout * C * k * k
    extra = ""
    if os.environ.get("PBE_GEMM_DEBUG"):
        import ctypes
        c = (ctypes.c_longlong * 8)()
        lib.pbe_debug_gemm_counters(c)
        if c[0] > 0:
            extra = f"  | CTA0 MMA warp: {c[0]} cyc, wait TMA {100*c[1]/c[0]:.0f}%, wait TMEM {100*c[2]/c[0]:.0f}%, {c[4]} k-iters, {c[0]/max(c[4],1):.0f} cyc/iter"
            if c[5] > 0:
                extra += f" | epi warp: {c[5]} cyc, wait acc {100*c[6]/c[5]:.0f}%, wait slot {100*(c[7]>>32)/c[5]:.0f}%, barrier {100*(c[7]&0xffffffff)/c[5]:.0f}%"
    print(f"{name:28s} {ms*1e3:8.1f} us {fl/ms/1e9:8.1f} TF/s{extra}", flush=True)

which = sys.argv[1:] or ["proj_out", "geglu", "conv2", "conv1", "ffout", "lowres"]
if "proj_out" in which: case("proj_out 64^2 320->320 +res", 16, 64, 64, 320, 1, 320, residual=True)
if "geglu" in which: case("geglu 64^2 320->2560", 16, 64, 64, 320, 1, 2560, mode=1, f32=False)
if "conv2" in which: case("conv2 64^2 320->320 +res", 16, 64, 64, 320, 3, 320, residual=True)
if "conv1" in which: case("conv1 64^2 320->320 +rb", 16, 64, 64, 320, 3, 320, rowbias=True)
if "ffout" in which: case("ff.out 64^2 1280->320 +res", 16, 64, 64, 1280, 1, 320, residual=True, f32=False, b16=True)
if "lowres" in which: case("conv 8^2 1280->1280 +res", 16, 8, 8, 1280, 3, 1280, residual=True)
if "nores" in which: case("proj 64^2 320->320 no res", 16, 64, 64, 320, 1, 320)
if "clip" in which:
    case("clip fc1 M=257 1024->4096", 1, 1, 257, 1024, 1, 4096, f32=False, b16=True)
    case("clip fc2 M=257 4096->1024 +res", 1, 1, 257, 4096, 1, 1024, residual=True)
    case("clip fc1 M=2056 1024->4096", 1, 1, 2056, 1024, 1, 4096, f32=False, b16=True)
    case("same as [8,1,257]", 8, 1, 257, 1024, 1, 4096, f32=False, b16=True)
if "vae" in which:
    case("vae L0 conv1 512^2 128->128", 8, 512, 512, 128, 3, 128)
    case("vae L0 conv2 512^2 128->128 +res", 8, 512, 512, 128, 3, 128, residual=True)
    case("vae L1 conv 256^2 256->256 +res", 8, 256, 256, 256, 3, 256, residual=True)
