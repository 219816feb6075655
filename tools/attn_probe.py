"""Time pbe_op_self_attention at one shape (default: the level-0 shape N=4096, 8 heads x 40, CFG batch 16); used for A/B
runs of the attention kernels (PBE_ATTN_KERNEL=2|3, PBE_ATTN_POLY=0|2|4|6|8, PBE_ATTN_RERUN=0|1, PBE_ATTN_TWO_PASS=1 --
read once per process) and under ncu.  Prints one line: us per call, algorithmic TFLOP/s, exponentials / ns / SM,
and the relative L2 error against the fp32 softmax on sample 0."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pbe_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N, heads, d = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (4096, 8, 40)
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 20
C = heads * d
g = torch.Generator().manual_seed(1)
qk = torch.randn(B, N, 2 * C, generator=g).to(dev).bfloat16()
Np = (N + 7) // 8 * 8
vt = torch.randn(B, C, Np, generator=g).to(dev).bfloat16()
out = torch.empty(B, N, C, device=dev, dtype=torch.float16 if lib.pbe_get_operand_format() else torch.bfloat16)
for _ in range(3):
    assert lib.pbe_op_self_attention(qk.data_ptr(), vt.data_ptr(), out.data_ptr(), B, N, heads, d, st) == 0, lib.pbe_last_error()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    lib.pbe_op_self_attention(qk.data_ptr(), vt.data_ptr(), out.data_ptr(), B, N, heads, d, st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
sp = lambda t: t.float().view(1, -1, heads, d).permute(0, 2, 1, 3)
q0, k0, v0 = qk[:1, :, :C], qk[:1, :, C:], vt[:1, :, :N].transpose(1, 2)
sim = torch.einsum("bhid,bhjd->bhij", sp(q0), sp(k0)) * d ** -0.5
ref = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), sp(v0)).permute(0, 2, 1, 3).reshape(1, N, C)
rel = ((out[:1].float() - ref).norm() / ref.norm()).item()
env = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("PBE_ATTN"))
print(f"attention B={B} N={N} heads={heads} d={d} [{env or 'defaults'}]: {ms * 1e3:.1f} us, {4.0 * B * N * N * C / ms / 1e9:.1f} TFLOP/s, "
      f"{B * heads * N * N / ms / 1e6 / 148:.2f} exp/ns/SM, rel-L2 {rel:.3e}")
