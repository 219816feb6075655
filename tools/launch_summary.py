"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of tools/ncu_target.py: per-kernel-family time of
the LAST U-Net call (= the last graph replay), as shares.  Usage: python tools/launch_summary.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
ci = {h: i for i, h in enumerate(rows[hi])}
data = [(r[ci["Kernel Name"]], float(r[ci["Metric Value"]])) for r in rows[hi + 1:] if len(r) > ci["Metric Value"]]
# the last U-Net call starts at the last launch of the sinusoidal-embedding kernel
starts = [i for i, (k, _) in enumerate(data) if "timestep_embedding" in k]
last = data[starts[-1]:] if starts else data


def family(k):
    for pat, name in (("conv_gemm", "conv_gemm_kernel (tcgen05 implicit GEMM)"), ("flash_attn2", "flash_attn2_kernel"),
                      ("flash_attn", "flash_attn_kernel"), ("splitk", "split-K reduce"), ("groupnorm", "GroupNorm kernels"),
                      ("gn_", "GroupNorm kernels"), ("layernorm", "LayerNorm"), ("upsample", "nearest-2x upsample")):
        if pat in k:
            return name
    return "other (pack / unpack / embeddings / GEMV / memcpy)"


agg = collections.OrderedDict()
for k, ns in last:
    a = agg.setdefault(family(k), [0, 0.0])
    a[0] += 1
    a[1] += ns
tot = sum(a[1] for a in agg.values())
print(f"ncu launch list (gpu__time_duration.sum, --clock-control none) of tools/ncu_target.py: one U-Net call, v1.yaml, "
      f"CFG batch 16, 64x64 latent")
print(f"last graph replay: {len(last)} kernels, {tot / 1e6:.2f} ms summed (cold-cache, serialised under ncu: shares, not absolutes)\n")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:56s} n={n:4d} {ns / 1e6:9.3f} ms {100 * ns / tot:6.1f} %")
