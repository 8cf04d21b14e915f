# round-2 ncu evidence (run on the GPU box through gpurun; every profiled command has exited 0 without ncu before)
set -x
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
timeout 300 $B > $O/r2_prof_plain.json 2> $O/r2_prof_plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file $O/r2_launches.csv $B > $O/r2_launches_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lstm_|l0_grad_gemm|gemm_tn_tma|gemm_nt_tma' -s 14 -c 11 -o $O/r2_headline_full $B > $O/r2_headline_full.log 2>&1
ncu -i $O/r2_headline_full.ncu-rep --page raw --csv > $O/r2_headline_full.raw.csv 2>/dev/null; rm -f $O/r2_headline_full.ncu-rep   # (gpurun_out travels back only below 64 MiB)
C5="python tools/bench_configs.py --config 5 --len 800 --batch 256"
timeout 200 $C5 > $O/r2_prof_c5_plain.json 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_config5.csv $C5 > $O/r2_launches_config5_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cltc -s 9 -c 6 -o $O/r2_cltc_full $C5 > $O/r2_cltc_full.log 2>&1
ncu -i $O/r2_cltc_full.ncu-rep --page raw --csv > $O/r2_cltc_full.raw.csv 2>/dev/null; rm -f $O/r2_cltc_full.ncu-rep
ls -la $O
