"""profiles/r2_* from the raw outputs of tools/make_profiles.sh in gpurun_out/: launch-list summaries, ncu summary table, traffic table.
usage: python tools/build_profiles.py"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3,
         "second": 1e6}


def load(f):
    rows = list(csv.reader(open(f)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def val(r, U, k):
    return float(r[k].replace(",", "")) * SCALE.get(U[k], 1)


def family(n):
    if "lstm_bwd_kernel<64, 1, 0, 0" in n: return "lstm_bwd_upper"
    if "lstm_bwd_kernel<64, 1, 0, 1" in n: return "lstm_bwd_l0"
    if "lstm_fwd_kernel<64, 1, 0, 1" in n: return "lstm_fwd_l0"
    if "lstm_fwd_kernel<64, 1, 0, 0" in n: return "lstm_fwd_upper"
    if "l0_grad_gemm" in n: return "l0_grads"
    if "gemm_tn_tma" in n: return "gemm_tn_dw"
    if "gemm_nt_tma_kernel<256" in n: return "gemm_nt_xproj"
    if "gemm_nt_tma_kernel<128" in n: return "gemm_nt_dgrad"
    return None


def line(r, U, label):
    d, rd, wr = val(r, U, "gpu__time_duration.sum"), val(r, U, "dram__bytes_read.sum"), val(r, U, "dram__bytes_write.sum")
    return (f"{label:66s} | {d:8.1f} | {rd / 1e6:8.1f} | {wr / 1e6:8.1f} | {float(r['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']):5.1f} | "
            f"{float(r['sm__warps_active.avg.pct_of_peak_sustained_active']):5.1f} | {float(r['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']):5.1f} | "
            f"{r['launch__registers_per_thread']} | {r['launch__grid_size']}"), d, rd + wr


def main():
    for src, dst, what in (("r2_launches.csv", "r2_launches_summary.txt",
                            "ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras (round 2, fp32 mode)"),
                           ("r2_launches_config5.csv", "r2_launches_config5_summary.txt",
                            "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python tools/bench_configs.py --config 5 --len 800 --batch 256 (round 2, fp32 mode: eval forward x2, then 2 training steps)")):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), os.path.join(O, src), what], capture_output=True, text=True).stdout
        open(os.path.join(P, dst), "w").write(out)
        with open(os.path.join(O, src)) as a, open(os.path.join(P, src), "w") as b:
            b.write(a.read())
    plain = json.loads(open(os.path.join(O, "r2_prof_plain.json")).read().strip().splitlines()[-1])
    rows_per_step = plain["run"]["token_rows_per_step"]
    hdr = "kernel | duration us | dram read MB | dram write MB | tensor pipe % | warps active % | dram % of peak | regs | grid"
    lines = ["# ncu --set full --clock-control none -k regex:lstm_|l0_grad_gemm|gemm_tn_tma|gemm_nt_tma -s 14 -c 11 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras",
             "# (round 2, fp32 mode, headline workload; one training step: 11 launches incl. the two phases of the layer-0 launches)", hdr]
    R, U = load(os.path.join(O, "r2_headline_full.raw.csv"))
    fam = {}
    for r in R:
        n = r["Kernel Name"]
        f = family(n)
        ln, d, b = line(r, U, f"{(f or '?'):15s} {n[20:70]}")
        lines.append(ln)
        if f:
            a = fam.setdefault(f, [0.0, 0.0, 0])
            a[0] += b
            a[1] += d
            a[2] += 1
    per_step = {"lstm_bwd_upper": 1, "lstm_bwd_l0": 2, "lstm_fwd_l0": 2, "lstm_fwd_upper": 1, "l0_grads": 1, "gemm_tn_dw": 1, "gemm_nt_xproj": 1, "gemm_nt_dgrad": 1}
    traffic = {"source": "ncu --set full --clock-control none -k regex:lstm_|l0_grad_gemm|gemm_tn_tma|gemm_nt_tma -s 14 -c 11 python bench.py --steps 2 --warmup 3 "
                         "--no-cpu-baseline --no-extras (tools/make_profiles.sh, round 2, fp32 mode); a two-phase family is the sum of its launches",
               "tokens_per_chain_at_capture": rows_per_step, "kernels": {}}
    for f, (b, d, n) in fam.items():
        reps = n / per_step[f]
        traffic["kernels"][f] = {"dram_bytes_per_launch": b / reps, "dram_bytes_per_token": b / reps / rows_per_step, "ncu_duration_us": d / reps}
    json.dump(traffic, open(os.path.join(P, "r2_traffic.json"), "w"), indent=1)
    lines += ["", "# tcgen05 cluster kernels: ncu --set full ... -k regex:cltc -s 9 -c 6 python tools/bench_configs.py --config 5 --len 800 --batch 256 (first training step: "
              "3 forward + 3 backward launches, tiles of 40, 14 clusters x 8 CTAs)", hdr]
    R, U = load(os.path.join(O, "r2_cltc_full.raw.csv"))
    for r in R:
        lines.append(line(r, U, r["Kernel Name"][20:86])[0])
    open(os.path.join(P, "r2_ncu_summary.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[-8:]))


if __name__ == "__main__":
    main()
