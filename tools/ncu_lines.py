"""Per-CUDA-source-line warp-sample shares from `ncu --page source --csv --print-source cuda` (usage: ncu_lines.py dump.csv kernel_index [min_share])."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
k = int(sys.argv[2])
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
secs, cur = [], None
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        secs.append(cur)
    elif cur is not None and cur["hdr"] is None and "# Samples" in r:
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None:
        cur["rows"].append(r)
print(len(secs), [s["name"][:70] for s in secs])
s = secs[k]
h = {n: i for i, n in enumerate(s["hdr"])}
stall_cols = [n for n in s["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
def val(r, n):
    try:
        return int(float(r[h[n]] or 0))
    except (ValueError, IndexError):
        return 0
tot = sum(val(r, "# Samples") for r in s["rows"])
print("kernel", s["name"][:100], "samples", tot)
for r in s["rows"]:
    n = val(r, "# Samples")
    if n > min_share * tot:
        why = sorted(((val(r, c), c[6:]) for c in stall_cols), reverse=True)[:2]
        print(f"{n:7d} {100 * n / tot:5.1f}%  {r[0]:>5} {r[h['Source']][:100]:100s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
