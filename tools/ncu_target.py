"""One U-Net call at the bench's configuration (v1.yaml U-Net, CFG batch 16, 64x64 latent) through the C ABI: the target
of the ncu launch-list pass (profiles/README.md).  The full bench command launches ~40k kernels per timed step, which ncu
cannot walk inside the GPU budget; this runs the same engine plan once eagerly (+ its graph capture and one replay)."""
import os, sys
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
ctx = torch.randn(Bc, 1, 768, generator=g).to(dev)
net.set_context(ctx)
out = torch.empty(Bc, 4, hw, hw, device=dev)
net.run(x, t, out=out)      # first call: eager warm pass + graph capture + replay
torch.cuda.synchronize()
net.run(x, t, out=out)      # steady state: one graph replay
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
