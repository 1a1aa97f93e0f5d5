mkdir -p gpurun_out
timeout 900 python bench.py --workload bunnies_30m --no-weak --no-cpu-baseline --no-parity --no-e2e --repeats 2 > gpurun_out/r2s_bunnies.json 2> gpurun_out/r2s_bench.err
