mkdir -p gpurun_out
for m in 1 0 1; do
MPM_G2P_TILE=$m timeout 200 python bench.py --workload multimat_12m --no-weak --no-cpu-baseline --no-parity --no-e2e --repeats 4 > gpurun_out/r2u_bench12_$m.json 2>> gpurun_out/r2u_bench.err
timeout 10 python tools/bench_brief.py < gpurun_out/r2u_bench12_$m.json
done
