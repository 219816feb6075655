"""Fit u(t) ~ log2(erfc(t)) on [0, 4] (weighted minimax by Lawson iterations) so that erf(t) = 1 - 2^u(t); prints the
coefficients of the GEGLU epilogue (pbe_b200/csrc/gemm_tc.cu geglu2), with 1/sqrt(2) folded in, and the fp32 errors."""
import numpy as np
from scipy.special import erf, erfc

DEG = 5
t = np.linspace(0, 4.0, 20001)
y = np.log2(erfc(t))
w = erfc(t) * np.log(2)            # d erf = ln2 * erfc * du
A = np.stack([t ** k for k in range(1, DEG + 1)], 1)
lw = np.ones_like(t)
for _ in range(60):
    W = w * lw
    c, *_ = np.linalg.lstsq(A * W[:, None], y * W, rcond=None)
    err = np.abs(w * (A @ c - y))
    lw = lw * (err / err.max() + 1e-3) ** 0.5
    lw /= lw.max()
cp = np.float32([c[k] * 2 ** (-(k + 1) / 2) for k in range(DEG)])
print("coefficients of |g|^1..5:", [f"{x:.9e}f" for x in cp])
g = np.linspace(-12, 12, 400001).astype(np.float32)
tt = np.minimum(np.abs(g), np.float32(4 * 2 ** 0.5))
u = np.full_like(tt, cp[4])
for k in (3, 2, 1, 0):
    u = u * tt + cp[k]
u = u * tt
h = np.copysign(np.float32(0.5) - np.float32(0.5) * np.exp2(u), g) + np.float32(0.5)
ref = g.astype(np.float64) * 0.5 * (1 + erf(g.astype(np.float64) / np.sqrt(2)))
print("max |gelu error|:", np.abs(g * h - ref).max())
