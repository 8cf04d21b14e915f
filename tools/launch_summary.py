"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file f.csv`): total time, share and launch count per kernel.
usage: python tools/launch_summary.py launches.csv ["header line"]"""
import collections
import csv
import sys

rows = []
with open(sys.argv[1], newline="") as fh:
    lines = [ln for ln in fh if not ln.startswith("==")]
rd = csv.reader(lines)
hdr = None
for r in rd:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = {n: i for i, n in enumerate(r)}
        continue
    if len(r) <= hdr["Metric Value"] or r[hdr["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[hdr["Metric Value"]].replace(",", ""))
    unit = r[hdr["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    rows.append((r[hdr["Kernel Name"]], us))
tot = sum(u for _, u in rows)
agg = collections.OrderedDict()
for k, u in rows:
    a = agg.setdefault(k, [0.0, 0])
    a[0] += u
    a[1] += 1
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print(f"# {len(rows)} launches, {tot:.1f} us in total (per-launch times are cold-cache / serialised under ncu: compare SHARES)")
for k, (u, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{u:10.1f} us {100 * u / tot:5.1f}%  n={n:3d}  {k[:170]}")
