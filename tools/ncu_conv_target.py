"""One launch shape for `ncu --set full`: the dominant kernel of a U-Net call at the bench configuration -- the first-level
3x3 convolution (CFG batch 16, 64x64 latent, 320 -> 320 channels, per-sample bias, 16-bit output), through the C ABI.
Eight identical launches; capture one with  -k regex:conv_gemm_kernel --launch-skip 5 --launch-count 1.
Algorithmic bytes of one launch: 16*64*64*320*2 (activations) + 9*320*320*2 (weights) + 16*64*64*320*2 (output) = 85.7 MB;
algorithmic FLOPs: 2 * 65536 * 320 * 2880 = 120.8 GFLOP."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pbe_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
Nb, H, W, C, Cout, k = 16, 64, 64, 320, 320, 3
f16 = bool(lib.pbe_get_operand_format())
dt = torch.float16 if f16 else torch.bfloat16
x = torch.randn(Nb, H, W, C, device=dev).to(dt)
w = (torch.randn(k * k, Cout, C, device=dev) / (C * k * k) ** 0.5).to(dt)
bias = torch.randn(Cout, device=dev)
rb = torch.randn(Nb, Cout, device=dev)
out = torch.empty(Nb, H, W, Cout, device=dev, dtype=dt)
p = _lib.ptr
for _ in range(8):
    rc = lib.pbe_op_conv_gemm(p(x), Nb, H, W, C, k, 1, p(w), Cout, 0, p(bias), p(rb), None, None, p(out), None, 0, 0, st)
    assert rc == 0, lib.pbe_last_error()
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
