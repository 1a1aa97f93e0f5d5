mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^Voxelizer\|^$" | tail -30 > gpurun_out/r2j_tests.log
timeout 900 python bench.py --no-weak --no-cpu-baseline --no-e2e > gpurun_out/r2j_bench.json 2>> gpurun_out/r2j_bench.err
timeout 600 python bench.py --workload cube_drop_4m --no-weak --no-cpu-baseline --no-parity --no-e2e > gpurun_out/r2j_bench4.json 2>> gpurun_out/r2j_bench.err
for m in 0 1 2 3; do
MPM_BENCH_MATERIAL=$m timeout 600 python bench.py --workload multimat_12m --no-weak --no-cpu-baseline --no-parity --no-e2e --repeats 6 > gpurun_out/r2j_mat$m.json 2>> gpurun_out/r2j_bench.err
done
