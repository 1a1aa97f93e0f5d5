mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "^Voxelizer\|^$" | tail -30 > gpurun_out/r2x_tests.log
tail -3 gpurun_out/r2x_tests.log
timeout 900 python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
timeout 10 python tools/bench_brief.py gpurun_out/r2x_bench.json
timeout 300 python bench.py --workload cube_drop_4m --no-weak --no-cpu-baseline --no-parity --no-e2e > gpurun_out/r2x_bench4.json 2>> gpurun_out/r2x_bench.err
timeout 10 python tools/bench_brief.py gpurun_out/r2x_bench4.json
timeout 600 python bench.py --impl reference > gpurun_out/r2x_ref.json 2>> gpurun_out/r2x_bench.err
tail -c 600 gpurun_out/r2x_ref.json
