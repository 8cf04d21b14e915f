"""Mean per-phase cycle counters printed by an IB200_PROF build (usage: prof_summary.py log)."""
import collections
import json
import re
import sys

mode, agg = None, collections.defaultdict(list)
for ln in open(sys.argv[1]):
    if ln.startswith("=="):
        mode = ln.split()[1]
        continue
    if ln.startswith("{"):
        d = json.loads(ln)
        print(mode, "infer", round(d["infer_ms"], 1), "train", round(d.get("train_ms", 0), 1),
              {k: v for k, v in d.get("train_kernel_ms", {}).items() if k.startswith("lstm")})
        continue
    m = re.match(r"(\w+ \w+)\s+\[(.*)\] T=(\d+) cycles/step: (.*)", ln)
    if m:
        agg[(mode, m.group(1), m.group(2))].append([int(x) for x in m.group(4).split()])
for k, v in agg.items():
    n = len(v)
    mean = [sum(c) // n for c in zip(*v)]
    names = k[2].split()
    print(k[0], k[1], " ".join(f"{a}={b}" for a, b in zip(names, mean)), "| sum", sum(mean), "n", n)
