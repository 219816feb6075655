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
# the sampler's fused update (K11) and input builder at the same batch: the HBM-bound elementwise kernels of the path
from pbe_b200 import _lib
import ctypes
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
lat = [torch.randn(B, 4, hw, hw, device=dev) for _ in range(9)]
x9 = torch.empty(Bc, 9, hw, hw, device=dev)
f = ctypes.c_float
for _ in range(3):
    assert lib.pbe_build_unet_input(lat[0].data_ptr(), lat[1].data_ptr(), lat[2][:, :1].contiguous().data_ptr(), x9.data_ptr(), B, 4, 4, 1,
                                    hw * hw, 2, st) == 0
    assert lib.pbe_sampler_step(out[:B].data_ptr(), out[B:].data_ptr(), f(5.0), 1, 3, lat[3].data_ptr(), lat[4].data_ptr(),
                                lat[5].data_ptr(), lat[0].data_ptr(), f(0.5), f(0.6), f(0.0), f(0.7), None, f(1.0), lat[6].data_ptr(),
                                lat[7].data_ptr(), lat[8].data_ptr(), lat[0].numel(), st) == 0
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
