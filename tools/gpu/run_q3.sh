mkdir -p gpurun_out
timeout 300 python bench.py --workload cube_drop_4m --quant --no-weak --no-cpu-baseline --no-parity --no-e2e > gpurun_out/r2q3_cube_quant.json 2>> gpurun_out/r2q3.err
timeout 10 python tools/bench_brief.py gpurun_out/r2q3_cube_quant.json
timeout 600 python bench.py --workload bunnies_30m --quant --no-weak --no-cpu-baseline --no-parity --no-e2e --repeats 2 > gpurun_out/r2q3_bunnies_quant.json 2>> gpurun_out/r2q3.err
timeout 10 python tools/bench_brief.py gpurun_out/r2q3_bunnies_quant.json
timeout 600 python bench.py --workload bunnies_30m --no-weak --no-cpu-baseline --no-parity --no-e2e --repeats 2 > gpurun_out/r2q3_bunnies.json 2>> gpurun_out/r2q3.err
timeout 10 python tools/bench_brief.py gpurun_out/r2q3_bunnies.json
tail -3 gpurun_out/r2q3.err
