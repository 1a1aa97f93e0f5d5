mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modes.py -x -q -m gpu > gpurun_out/r2q2_modes.log 2>&1
tail -25 gpurun_out/r2q2_modes.log
timeout 200 python bench.py --workload multimat_12m --no-weak --no-cpu-baseline --no-parity --no-e2e --repeats 4 > gpurun_out/r2q2_bench12.json 2>> gpurun_out/r2q2_bench.err
timeout 10 python tools/bench_brief.py gpurun_out/r2q2_bench12.json
