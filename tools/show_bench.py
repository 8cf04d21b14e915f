"""Print ms/step and the per-family kernel times of bench JSON lines: python tools/show_bench.py gpurun_out/bench12*.json"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(f, "ERR", e)
        continue
    k = d.get("kernel_families", {})
    print(f"{f}: {d['ms_per_step']:.3f} ms/step  {d['value']:.0f} {d['unit']}  e2e {d.get('e2e', {}).get('value', 0):.0f}")
    print("   ", {n: round(v["ms_per_step"], 3) for n, v in k.items() if v["ms_per_step"] > 0.05})
