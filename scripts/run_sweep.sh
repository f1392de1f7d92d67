# GPU-box helper: parity of the parameter-set kernels, then the per-set throughput sweep (arguments: pytest -k filter, sweep list)
timeout 500 python -m pytest tests/test_gpu_param_sets.py -m gpu -q -x -k "${1:-test}" 2>&1 | tail -4 > gpurun_out/param_sets_last.log; tail -4 gpurun_out/param_sets_last.log
timeout 300 python bench.py --param-sweep "${2:-1_0,1_1,2_0,1_2,2_3,1_4,3_3,4_3,4_4}" --cpu-sample 0 > gpurun_out/param_sweep_last.json 2> gpurun_out/param_sweep_last.err; tail -3 gpurun_out/param_sweep_last.err
python scripts/print_sweep.py gpurun_out/param_sweep_last.json
