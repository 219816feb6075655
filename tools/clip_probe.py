"""Time FrozenCLIPImageEmbedder.forward (ViT-L/14 tower + mapper + final_ln) through the C ABI."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import clip_ref as K          # synthetic weights / images only
from pbe_b200.clip import FrozenCLIPImageEmbedder
from pbe_b200 import _lib

dev = torch.device("cuda:0")
cfg = K.V1_CLIP_CFG
m = FrozenCLIPImageEmbedder(**cfg)
m.load_state_dict(K.make_state_dict(cfg, 321), strict=True)
m = m.to(dev).eval()
for B in [int(a) for a in sys.argv[1:]] or [1, 8]:
    x = K.synthetic_exemplars(B, 224, seed=3).to(dev)
    for _ in range(3):
        z = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        z = m(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = B * 2 * (257 * (24 * (4 * 1024 * 1024 + 2 * 1024 * 4096) + 0) + 256 * 588 * 1024 + 24 * 2 * 257 * 257 * 1024)
    print(f"CLIP front-end B={B}: {ms:.3f} ms per batch, {ms/B:.3f} ms per image, ~{flops/ms/1e9:.0f} TFLOP/s, "
          f"{_lib.load().pbe_clip_launches_per_encode(m._engine)} kernels")
