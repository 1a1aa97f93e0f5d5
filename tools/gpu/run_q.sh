mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_modes.py tests/test_gpu_api.py -m gpu -q 2>&1 | grep -v "^Voxelizer\|^$" | tail -40 > gpurun_out/r2q_tests.log
timeout 600 python bench.py --workload cube_drop_4m --g2p2g --quant --no-weak --no-cpu-baseline --no-parity --no-e2e > gpurun_out/r2q_g2p2g4.json 2>> gpurun_out/r2q_bench.err
timeout 600 python bench.py --workload multimat_12m --g2p2g --quant --no-weak --no-cpu-baseline --no-parity --no-e2e --repeats 4 > gpurun_out/r2q_g2p2g12.json 2>> gpurun_out/r2q_bench.err
