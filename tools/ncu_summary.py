"""Summarise an `ncu --page source --csv --print-source sass` dump: per kernel, stall-reason totals and the hottest instructions.
usage: python tools/ncu_summary.py src.csv [top_n] [kernel_index ...]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
only = set(int(x) for x in sys.argv[3:])
sections, cur = [], None
for row in csv.reader(open(path)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        sections.append(cur)
    elif row[0] == "Address":
        cur["hdr"] = row
    elif cur is not None and cur["hdr"] is not None:
        cur["rows"].append(row)
for k, s in enumerate(sections):
    if only and k not in only:
        continue
    h = {n: i for i, n in enumerate(s["hdr"])}
    stall_cols = [n for n in s["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
    tot = {n: 0 for n in stall_cols}
    nsamp, ninst = 0, 0
    for r in s["rows"]:
        for n in stall_cols:
            try:
                tot[n] += int(float(r[h[n]] or 0))
            except ValueError:
                pass
        nsamp += int(float(r[h["# Samples"]] or 0))
        ninst += int(float(r[h["Instructions Executed"]] or 0))
    print(f"=== [{k}] {s['name'][:90]}  samples={nsamp} warp-instr={ninst}")
    print("   " + "  ".join(f"{n[6:]}={v} ({100*v/max(nsamp,1):.0f}%)" for n, v in sorted(tot.items(), key=lambda x: -x[1]) if v > 0.01 * nsamp))
    rows = sorted(s["rows"], key=lambda r: -int(float(r[h["# Samples"]] or 0)))[:top]
    for r in rows:
        why = sorted(((int(float(r[h[n]] or 0)), n[6:]) for n in stall_cols), reverse=True)[:2]
        print(f"   {int(float(r[h['# Samples']])):7d}  {r[h['Source']][:70]:70s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
