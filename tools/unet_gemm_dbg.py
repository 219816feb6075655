"""In-situ wait counters of every conv_gemm op of one U-Net call (PBE_GEMM_DEBUG, Engine::profile_forward): which role of
CTA 0 waits on what, per layer.  Usage: python tools/unet_gemm_dbg.py [B] [hw]  (stderr carries the [gemm-dbg] lines)."""
import os, sys
os.environ["PBE_GEMM_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import unet_ref as U          # synthetic weights only
from pbe_b200.unet import UNetModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda:0")
cfg = U.V1_CFG
net = UNetModel(**cfg)
net.load_state_dict(U.make_state_dict(cfg, 321), strict=False)
net = net.to(dev).eval()
Bc = 2 * B
g = torch.Generator().manual_seed(1)
x = torch.randn(Bc, 9, hw, hw, generator=g).to(dev)
t = torch.full((Bc,), 501, dtype=torch.int64, device=dev)
net.set_context(torch.randn(Bc, 1, 768, generator=g).to(dev))
net.profile(x, t)
sys.stderr.write("==== second pass ====\n")
rows = net.profile(x, t)
for r in rows:
    if r["family"] == "conv_gemm":
        print(f"{r['name']:36s} {r['ms']*1e3:8.1f} us")
