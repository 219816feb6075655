"""Pick the launch to profile from an ncu launch list (tools/ncu_target.py under `--metrics gpu__time_duration.sum --csv`):
prints the 0-based invocation index, AMONG THE LAUNCHES WHOSE NAME MATCHES `regex`, of the longest matching launch inside
the last U-Net call (the steady-state graph replay).  `select` (optional) narrows which of those launches may be picked
(ncu's --kernel-id counts invocations by BASE name, the CSV shows template arguments).
Usage: python tools/ncu_pick.py launches.csv <base-name regex> [select regex] [nth-longest]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
pat = re.compile(sys.argv[2])
sel = re.compile(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].isdigit() else pat
nth = int(sys.argv[-1]) if len(sys.argv) > 3 and sys.argv[-1].isdigit() else 0
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
ci = {h: i for i, h in enumerate(rows[hi])}
data = [(r[ci["Kernel Name"]], float(r[ci["Metric Value"]])) for r in rows[hi + 1:] if len(r) > ci["Metric Value"]]
starts = [i for i, (k, _) in enumerate(data) if "timestep_embedding" in k]
first = starts[-1] if starts else 0
match_idx = [i for i, (k, _) in enumerate(data) if pat.search(k)]
inside = sorted((i for i in match_idx if i >= first and sel.search(data[i][0])), key=lambda i: -data[i][1])
if not inside:
    sys.exit(f"no launch matching {sys.argv[2]!r} in the last U-Net call")
pick = inside[min(nth, len(inside) - 1)]
print(match_idx.index(pick))
