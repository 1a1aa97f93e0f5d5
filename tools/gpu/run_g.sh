set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^Voxelizer\|^$" | tail -30 > gpurun_out/r2g_tests.log
for d in 0 1; do
MPM_DEFER_SVD=$d timeout 600 python bench.py --workload multimat_12m --no-weak --no-cpu-baseline --no-parity --no-e2e > gpurun_out/r2g_bench12_d$d.json 2>> gpurun_out/r2g_bench.err
MPM_DEFER_SVD=$d timeout 600 python bench.py --workload cube_drop_4m --no-weak --no-cpu-baseline --no-parity --no-e2e > gpurun_out/r2g_bench4_d$d.json 2>> gpurun_out/r2g_bench.err
done
timeout 900 python bench.py --no-weak --no-cpu-baseline --no-e2e > gpurun_out/r2g_bench.json 2>> gpurun_out/r2g_bench.err
