#!/bin/bash
# per-kernel durations of the optimizer kernels (ncu replays serialise the launches and run them cold: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'r21_|adamw' -c 60 --csv --log-file gpurun_out/r2_ncu_optim.csv \
  python tools/bench_optim.py > gpurun_out/r2_ncu_optim.log 2>&1
python tools/launch_summary.py gpurun_out/r2_ncu_optim.csv "ncu launch list of tools/bench_optim.py (optimizer kernels only)" | tee gpurun_out/r2_ncu_optim_summary.txt
