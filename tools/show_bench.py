import json, sys
d = json.load(open(sys.argv[1]))
print(f"value {d['value']:.0f} {d['unit']}  ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']:.0f} ({d['e2e']['ms_per_step']:.3f} ms)  launches {d['gpu_launches']}  clocks {d['clocks']}")
for k, v in d["kernel_families"].items():
    print(f"  {k:16s} {v['ms_per_step']:.3f} ms  x{v['launches_per_step']:.0f}  {100*v['share']:.1f}%")
print(" other:", d.get("other_variants_1gpu"))
r = d["roofline"]; print(" roofline:", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k != 'note'})
if "cpu_baseline" in d: print(" cpu:", d["cpu_baseline"])
