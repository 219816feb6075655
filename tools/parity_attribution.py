"""Which bf16 rounding point contributes how much of the ~9e-3 eps error?  ("parity headroom".)

Runs the fp32 oracle U-Net (CPU) with ONE class of operands rounded to bf16 at a time -- the product's rounding points,
oracle/bf16_emul.py -- and prints the relative L2 error of eps against the all-fp32 result.  Independent error sources add
in quadrature, so the squares show the budget.  Development tool (imports the oracle; not part of the product).

    python tools/parity_attribution.py [v1|small] [t]
"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import sampler_ref as S, unet_ref as M

which = sys.argv[1] if len(sys.argv) > 1 else "v1"
tval = int(sys.argv[2]) if len(sys.argv) > 2 else 981
cfg = M.V1_CFG if which == "v1" else M.SMALL_CFG
hw = 64 if which == "v1" else 32
sd = M.make_state_dict(cfg, 321)
req = S.synthetic_request(1, hw, hw, seed=321)
x9 = torch.cat((req["x_T"], req["z_inpaint"], req["mask"]), 1)
x_in, c_in = torch.cat([x9] * 2), torch.cat((req["uc"], req["c"]))
t = torch.full((2,), tval, dtype=torch.int64)
r = lambda v: v.bfloat16().float()
of_conv, of_lin, of_einsum = F.conv2d, F.linear, torch.einsum


def run(flags):
    def conv2d(x, w, b=None, **kw):
        k = w.shape[-1]
        key = "conv3" if k == 3 else "conv1"
        xa = r(x) if key + "_act" in flags else x
        wa = r(w) if key + "_wt" in flags else w
        return of_conv(xa, wa, b, **kw)

    def linear(x, w, b=None):
        if x.dim() == 2:
            if w.shape[0] != w.shape[1] or True:
                pass
            if "emb" in flags and w.shape[1] == 4 * cfg["model_channels"] and w.shape[0] != 4 * cfg["model_channels"]:
                return of_lin(r(x), r(w), b)      # emb_layers run as one bf16 tensor-core GEMM in the engine
            return of_lin(x, w, b)
        if x.dim() == 3 and x.shape[1] == 1:
            return of_lin(x, w, b)
        xa = r(x) if "lin_act" in flags else x
        wa = r(w) if "lin_wt" in flags else w
        return of_lin(xa, wa, b)

    def einsum(eq, a, b):
        if eq == "bid,bjd->bij":
            return of_einsum(eq, r(a) if "qk" in flags else a, r(b) if "qk" in flags else b)
        return of_einsum(eq, r(a) if "p" in flags else a, r(b) if "v" in flags else b)

    M.F.conv2d, M.F.linear, M.torch.einsum = conv2d, linear, einsum
    try:
        with torch.no_grad():
            return M.unet_forward(sd, cfg, x_in, t, c_in)
    finally:
        M.F.conv2d, M.F.linear, M.torch.einsum = of_conv, of_lin, of_einsum


rel = lambda a, b: ((a - b).norm() / b.norm()).item()
t0 = time.time()
ref = run(set())
print(f"fp32 reference: {time.time() - t0:.1f} s", flush=True)
ALL = ["conv3_act", "conv3_wt", "conv1_act", "conv1_wt", "lin_act", "lin_wt", "qk", "p", "v", "emb"]
tot = 0.0
for f in ALL:
    e = rel(run({f}), ref)
    tot += e * e
    print(f"only {f:10s}: rel-L2 {e:.3e}   (squared share {e * e:.3e})", flush=True)
print(f"root of the sum of squares: {tot ** 0.5:.3e}")
print(f"all together              : {rel(run(set(ALL)), ref):.3e}")
