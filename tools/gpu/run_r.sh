mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "^Voxelizer\|^$" | tail -30 > gpurun_out/r2r_tests.log
timeout 600 python bench.py --workload multimat_12m --no-weak --no-cpu-baseline --no-parity --no-e2e --repeats 4 > gpurun_out/r2r_bench12.json 2>> gpurun_out/r2r_bench.err
timeout 900 python bench.py --no-weak --no-cpu-baseline --no-e2e > gpurun_out/r2r_bench.json 2>> gpurun_out/r2r_bench.err
timeout 600 python bench.py --workload cube_drop_4m --no-weak --no-cpu-baseline --no-parity --no-e2e > gpurun_out/r2r_bench4.json 2>> gpurun_out/r2r_bench.err
