set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ks_tc_kernel -s 1 -c 1 -o gpurun_out/r2_prof_ks_tc -f python scripts/ab_wide.py > gpurun_out/r2_prof_ks_tc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pbs_classic_kernel_v8x2 -s 2 -c 1 -o gpurun_out/r2_prof_v8x2_final -f python scripts/narrow_level_case.py 74 > gpurun_out/r2_prof_v8x2_final.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pbs_multibit_kernel_v8x2 -s 1 -c 1 -o gpurun_out/r2_prof_mb8x2_final -f python scripts/profile_case.py 2_2_g3 74 > gpurun_out/r2_prof_mb8x2_final.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2_ncu_bench.log 2>&1
tail -2 gpurun_out/r2_ncu_bench.log | cut -c1-300
ls -la gpurun_out/r2_prof_*final* gpurun_out/r2_prof_ks_tc* gpurun_out/r2_launches.csv
