"""Diagnose pbe_op_self_attention against the fp32 softmax reference: NaN positions and error by row."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pbe_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
B, N, heads, d = 2, 4096, 8, 40
if len(sys.argv) > 4:
    B, N, heads, d = map(int, sys.argv[1:5])
C = heads * d
g = torch.Generator().manual_seed(100 + N + d)
q = torch.randn(B, N, C, generator=g).to(dev).bfloat16()
k = torch.randn(B, N, C, generator=g).to(dev).bfloat16()
v = torch.randn(B, N, C, generator=g).to(dev).bfloat16()
qk = torch.cat((q, k), dim=-1).contiguous()
Np = (N + 7) // 8 * 8
vt = torch.zeros(B, C, Np, device=dev, dtype=torch.bfloat16)
vt[:, :, :N] = v.transpose(1, 2)
out = torch.zeros(B, N, C, device=dev, dtype=torch.float16 if lib.pbe_get_operand_format() else torch.bfloat16)
assert lib.pbe_op_self_attention(qk.data_ptr(), vt.data_ptr(), out.data_ptr(), B, N, heads, d, st) == 0
torch.cuda.synchronize()
sp = lambda t: t.float().view(B, N, heads, d).permute(0, 2, 1, 3)
sim = torch.einsum("bhid,bhjd->bhij", sp(q), sp(k)) * d ** -0.5
ref = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), sp(v))
o = sp(out)
nan = torch.isnan(o).any(-1)                      # [B, heads, N]
print("rows with NaN:", int(nan.sum()), "of", nan.numel())
if nan.any():
    idx = nan.nonzero()[:12].tolist()
    print("first NaN rows (b, head, row):", idx)
    print("NaN rows per (b, head):", nan.sum(-1).tolist())
    b0, h0, r0 = idx[0]
    s = sim[b0, h0, r0] * 1.4426950408889634
    tm = s.view(-1, 128).amax(-1)
    print("row", idx[0], "tile maxima (log2 units):", [round(x, 2) for x in tm.tolist()])
    print("tile log2-sums:", [round(x, 2) for x in torch.log2(torch.exp2(s.view(-1, 128)).sum(-1)).tolist()])
    print("out row:", o[b0, h0, r0, :8].tolist())
ok = ~nan
err = (o - ref).norm(dim=-1) / ref.norm(dim=-1)
print("rel-L2 over clean rows: mean %.3e max %.3e" % (err[ok].mean().item(), err[ok].max().item()))
bad = (err > 0.05) & ok
print("clean rows with rel err > 5e-2:", int(bad.sum()), bad.nonzero()[:8].tolist())
print("overall rel:", ((o - ref)[ok].norm() / ref[ok].norm()).item())
