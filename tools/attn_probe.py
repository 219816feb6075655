"""Run the level-0 self-attention shape (N=4096, 8 heads x 40) through pbe_op_self_attention; used under ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pbe_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N, heads, d = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (4096, 8, 40)
C = heads * d
qk = torch.randn(B, N, 2 * C, device=dev).bfloat16()
vt = torch.randn(B, C, N, device=dev).bfloat16()
out = torch.empty(B, N, C, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    assert lib.pbe_op_self_attention(qk.data_ptr(), vt.data_ptr(), out.data_ptr(), B, N, heads, d, st) == 0
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    lib.pbe_op_self_attention(qk.data_ptr(), vt.data_ptr(), out.data_ptr(), B, N, heads, d, st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"attention B={B} N={N} d={d}: {ms*1e3:.1f} us, {4.0*B*N*N*C/ms/1e9:.1f} TF/s, {B*heads*N*N/ms/1e6/148:.2f} exp/ns/SM")
