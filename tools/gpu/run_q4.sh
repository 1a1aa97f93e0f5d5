mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modes.py tests/test_gpu_substep.py -x -q -m gpu > gpurun_out/r2q4_modes.log 2>&1
tail -4 gpurun_out/r2q4_modes.log
timeout 300 python bench.py --workload cube_drop_4m --quant --no-weak --no-cpu-baseline --no-parity --no-e2e > gpurun_out/r2q4_cube_quant.json 2>> gpurun_out/r2q4.err
timeout 10 python tools/bench_brief.py gpurun_out/r2q4_cube_quant.json
timeout 300 python bench.py --workload cube_drop_4m --quant --g2p2g --no-weak --no-cpu-baseline --no-parity --no-e2e > gpurun_out/r2q4_cube_quant_g2p2g.json 2>> gpurun_out/r2q4.err
timeout 10 python tools/bench_brief.py gpurun_out/r2q4_cube_quant_g2p2g.json
timeout 300 python bench.py --workload cube_drop_4m --no-weak --no-cpu-baseline --no-parity --no-e2e > gpurun_out/r2q4_cube.json 2>> gpurun_out/r2q4.err
timeout 10 python tools/bench_brief.py gpurun_out/r2q4_cube.json
timeout 600 python bench.py --workload bunnies_30m --quant --no-weak --no-cpu-baseline --no-parity --no-e2e --repeats 2 > gpurun_out/r2q4_bunnies_quant.json 2>> gpurun_out/r2q4.err
timeout 10 python tools/bench_brief.py gpurun_out/r2q4_bunnies_quant.json
tail -3 gpurun_out/r2q4.err
