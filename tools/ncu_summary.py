"""Summarise an .ncu-rep (first kernel): key raw metrics + stall samples by opcode + hottest instructions."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
traffic_out = sys.argv[2] if len(sys.argv) > 2 else None     # optional: write {"dram_bytes": read + write, ...} (bench.py reads profiles/ncu_traffic.json)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for i, h in enumerate(hdr):
    if h in want or "issue_stalled" in h and "per_issue" not in h and False:
        print(f"{h:80s} {vals[i]} {rows[1][i]}")
if traffic_out:
    import json
    col = {h: i for i, h in enumerate(hdr)}
    def val(name):
        v, u = float(vals[col[name]]), rows[1][col[name]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    name_col = col.get("Kernel Name")
    json.dump({"dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
               "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
               "kernel": vals[name_col] if name_col is not None else None,
               "duration_us_under_ncu": float(vals[col["gpu__time_duration.sum"]]) * {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(rows[1][col["gpu__time_duration.sum"]], 1),
               "algorithmic_bytes": 85.7e6,
               "note": "DRAM read + write bytes of ONE launch of the dominant kernel -- the first-level 3x3 convolution of a U-Net "
                       "call at CFG batch 16, 64x64 (320 -> 320 channels, per-sample bias, 16-bit output; tools/ncu_conv_target.py) "
                       "-- from `ncu --set full` of the current build; algorithmic bytes of that launch: 85.7 MB (the 42 MB output "
                       "is largely still in L2 when the kernel ends); source: " + rep.split("/")[-1]}, open(traffic_out, "w"), indent=1)
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(vals[i] or 0) > 0.15:
        print(f"{h:80s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ci["# Samples"]]) for r in data)
agg, ex = collections.Counter(), collections.Counter()
for r in data:
    toks = r[ci["Source"]].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    agg[op] += int(r[ci["# Samples"]]); ex[op] += int(r[ci["Instructions Executed"]])
print("total samples", tot, "total inst", sum(ex.values()))
for op, c in agg.most_common(14):
    print(f"  {op:32s} {100*c/tot:5.1f}% samples   executed {ex[op]}")
for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:12]:
    st = {k[6:]: int(r[ci[k]]) for k in ci if k.startswith("stall_") and "Not" not in k and int(r[ci[k]]) > int(r[ci["# Samples"]]) * 0.15}
    print("  ", r[ci["# Samples"]], r[ci["Source"]].strip()[:64], st)
