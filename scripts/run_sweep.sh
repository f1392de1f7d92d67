timeout 500 python -m pytest tests/test_gpu_param_sets.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/param_sets3.log; tail -4 gpurun_out/param_sets3.log
timeout 300 python bench.py --param-sweep 1_0,1_1,2_0,1_2,2_3,1_4,3_3,4_3,4_4 --cpu-sample 0 > gpurun_out/param_sweep4.json 2> gpurun_out/param_sweep4.err; tail -3 gpurun_out/param_sweep4.err
python scripts/print_sweep.py gpurun_out/param_sweep4.json
