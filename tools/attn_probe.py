"""Run the level-0 self-attention shape (N=4096, 8 heads x 40) through pbe_op_self_attention; used under ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pbe_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N, heads, d = 4096, 8, 40
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

if os.environ.get("PBE_ATTN_TRACE"):
    import numpy as np
    attn2 = bool(os.environ.get("PBE_ATTN2"))
    T = N // 128 if attn2 else N // 64
    buf = torch.zeros(8 * T, dtype=torch.int64, device=dev)
    lib.pbe_debug_set_attention_trace(buf.data_ptr())
    lib.pbe_op_self_attention(qk.data_ptr(), vt.data_ptr(), out.data_ptr(), B, N, heads, d, st)
    torch.cuda.synchronize()
    lib.pbe_debug_set_attention_trace(None)
    t = buf.cpu().numpy().reshape(T, 8)
    names = ["wait s_full", "pass1(max)", "max exchange", "pv_done+corr", "pass2(exp)", "fence+arrive"] if attn2 else \
        ["wait s_full", "LDTM x2 + wait", "max+exp(+corr)", "wait pv_done(j-2)", "sum+pack+STS", "fence.proxy.async", "tc fence+arrive"]
    d_ = np.diff(t[:, :7] if attn2 else t[:, :8], axis=1)
    print("per-tile cycles (median over tiles 2..): " + ", ".join(f"{n}={int(np.median(d_[2:, i]))}" for i, n in enumerate(names)))
    print("tile period (median):", int(np.median(np.diff(t[2:, 0]))))
