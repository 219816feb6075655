"""First-contact check of the tcgen05 implicit-GEMM kernel against torch fp32 on the GPU. Prints per-case errors."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from pbe_b200 import _lib

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
lib = _lib.load()
dev = torch.device("cuda:0")


def run_case(name, Nb, H, W, C, ksize, stride, Cout, block_n=0, use_bias=True, use_rowbias=False, use_res=False,
             timing=False):
    g = torch.Generator(device="cpu").manual_seed(hash(name) % (2**31))
    x = torch.randn(Nb, H, W, C, generator=g).to(dev).bfloat16()
    w = (torch.randn(Cout, C, ksize, ksize, generator=g) / (C * ksize * ksize) ** 0.5).to(dev).bfloat16()
    wt = w.permute(2, 3, 0, 1).contiguous().view(ksize * ksize, Cout, C)
    bias = torch.randn(Cout, generator=g).to(dev) if use_bias else None
    Ho, Wo = H // stride, W // stride
    rowbias = torch.randn(Nb, Cout, generator=g).to(dev) if use_rowbias else None
    res = torch.randn(Nb, Ho, Wo, Cout, generator=g).to(dev) if use_res else None
    out = torch.full((Nb, Ho, Wo, Cout), float("nan"), device=dev)
    outb = torch.zeros((Nb, Ho, Wo, Cout), device=dev, dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.pbe_op_conv_gemm(x.data_ptr(), Nb, H, W, C, ksize, stride, wt.data_ptr(), Cout, 0,
                              _lib.ptr(bias), _lib.ptr(rowbias), _lib.ptr(res), out.data_ptr(), outb.data_ptr(), None,
                              0, block_n, st)
    if rc != 0:
        print(f"[{name}] rc={rc} err={lib.pbe_last_error().decode()}")
        return False
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, stride=stride, padding=ksize // 2)
    ref = ref.permute(0, 2, 3, 1)
    if rowbias is not None:
        ref = ref + rowbias[:, None, None, :]
    if res is not None:
        ref = ref + res
    err = (out - ref).abs().max().item()
    rel = ((out - ref).norm() / ref.norm()).item()
    errb = (outb.float() - ref).abs().max().item()
    nan = torch.isnan(out).sum().item()
    ok = rel < 2e-3 and nan == 0
    msg = f"[{name}] Nb={Nb} H={H} W={W} C={C} k={ksize} s={stride} Cout={Cout} bn={block_n}: max_abs={err:.3e} rel_l2={rel:.3e} bf16out_max={errb:.3e} nan={nan} {'OK' if ok else 'FAIL'}"
    if timing:
        for _ in range(3):
            lib.pbe_op_conv_gemm(x.data_ptr(), Nb, H, W, C, ksize, stride, wt.data_ptr(), Cout, 0, _lib.ptr(bias), None, None, out.data_ptr(), None, None, 0, block_n, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iters = 20
        for _ in range(iters):
            lib.pbe_op_conv_gemm(x.data_ptr(), Nb, H, W, C, ksize, stride, wt.data_ptr(), Cout, 0, _lib.ptr(bias), None, None, out.data_ptr(), None, None, 0, block_n, st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = 2.0 * Nb * Ho * Wo * Cout * C * ksize * ksize
        msg += f"  {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s"
    print(msg, flush=True)
    if not ok:
        d = (out - ref).abs()
        idx = torch.nonzero(d > 10 * max(err * 0.1, 1e-2))[:8]
        print("   sample bad idx:", idx.tolist())
        print("   out[0,0,0,:8]=", out[0, 0, 0, :8].tolist())
        print("   ref[0,0,0,:8]=", ref[0, 0, 0, :8].tolist())
    return ok


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    allok = True
    # plain GEMMs
    allok &= run_case("gemm_1chunk", 1, 1, 256, 64, 1, 1, 128, block_n=128)
    allok &= run_case("gemm_k256", 1, 1, 512, 256, 1, 1, 128, block_n=128)
    allok &= run_case("gemm_bn64", 1, 1, 512, 256, 1, 1, 64, block_n=64)
    allok &= run_case("gemm_bn160", 1, 1, 512, 320, 1, 1, 320, block_n=160)
    # (BLOCK_N = 256 is the GEGLU tile; a plain epilogue on it is refused by build_gemm_plan)
    allok &= run_case("gemm_bn32", 1, 1, 300, 320, 1, 1, 16, block_n=32)
    allok &= run_case("gemm_ragged_m", 1, 1, 1000, 128, 1, 1, 128, block_n=128, use_rowbias=True, use_res=True)
    # convs
    allok &= run_case("conv3_64", 2, 64, 64, 64, 3, 1, 128)
    allok &= run_case("conv3_32", 2, 32, 32, 128, 3, 1, 160, use_rowbias=True, use_res=True)
    allok &= run_case("conv3_16", 2, 16, 16, 128, 3, 1, 128)
    allok &= run_case("conv3_8", 3, 8, 8, 128, 3, 1, 128)
    allok &= run_case("conv3_4", 2, 4, 4, 64, 3, 1, 64)
    allok &= run_case("conv3_2", 2, 2, 2, 64, 3, 1, 64)
    allok &= run_case("conv3_96", 1, 96, 96, 64, 3, 1, 128)
    allok &= run_case("conv3_12", 3, 12, 12, 64, 3, 1, 128)
    allok &= run_case("conv3_s2_64", 2, 64, 64, 64, 3, 2, 128)
    allok &= run_case("conv3_s2_16", 2, 16, 16, 128, 3, 2, 128)
    allok &= run_case("conv1_32", 2, 32, 32, 192, 1, 1, 128)
    # perf probes (real layer shapes at Bc=16)
    run_case("perf_conv320_64", 16, 64, 64, 320, 3, 1, 320, block_n=160, timing=True)
    run_case("perf_conv320_64_bn128", 16, 64, 64, 320, 3, 1, 384, block_n=128, timing=True)
    run_case("perf_conv320_64_bn256", 16, 64, 64, 320, 3, 1, 512, block_n=256, timing=True)
    run_case("perf_conv640_32", 16, 32, 32, 640, 3, 1, 640, block_n=160, timing=True)
    run_case("perf_conv1280_16", 16, 16, 16, 1280, 3, 1, 1280, block_n=160, timing=True)
    run_case("perf_conv1280_8", 16, 8, 8, 1280, 3, 1, 1280, block_n=160, timing=True)
    run_case("perf_lin_ff", 1, 1, 65536, 1280, 1, 1, 320, block_n=160, timing=True)
    print("ALL OK" if allok else "SOME FAILED")
    sys.exit(0 if allok else 1)
