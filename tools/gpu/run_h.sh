set -x
mkdir -p gpurun_out
for d in 0 1; do
MPM_DEFER_SVD=$d timeout 600 python bench.py --workload multimat_12m --no-weak --no-cpu-baseline --no-parity --no-e2e > gpurun_out/r2h_bench12_d$d.json 2>> gpurun_out/r2h_bench.err
MPM_DEFER_SVD=$d timeout 600 python bench.py --workload cube_drop_4m --no-weak --no-cpu-baseline --no-parity --no-e2e > gpurun_out/r2h_bench4_d$d.json 2>> gpurun_out/r2h_bench.err
done
MPM_DEFER_SVD=1 timeout 900 python -m pytest tests/test_gpu_substep.py tests/test_gpu_parity2.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r2h_tests.log
