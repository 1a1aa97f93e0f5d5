set -x
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02g_launches_12m.csv python tools/profile_bench.py > gpurun_out/r02g_prof1.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -s 9 -c 9 -o gpurun_out/r02g_prof python tools/profile_bench.py > gpurun_out/r02g_prof2.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_p2g3|k_g2p" -s 2 -c 2 -o gpurun_out/r02g_prof_100m python tools/profile_bench.py --workload multimat_100m > gpurun_out/r02g_prof3.log 2>&1
timeout 300 python bench.py --workload cube_drop_4m --no-weak --no-cpu-baseline --no-parity > gpurun_out/r02g_bench4.json 2>> gpurun_out/r02g.err
timeout 10 python tools/bench_brief.py gpurun_out/r02g_bench4.json
timeout 600 python bench.py --workload bunnies_30m --no-weak --no-cpu-baseline --no-parity --no-e2e --repeats 2 > gpurun_out/r02g_bunnies.json 2>> gpurun_out/r02g.err
timeout 10 python tools/bench_brief.py gpurun_out/r02g_bunnies.json
timeout 600 python bench.py --workload bunnies_30m --quant --no-weak --no-cpu-baseline --no-parity --no-e2e --repeats 2 > gpurun_out/r02g_bunnies_q.json 2>> gpurun_out/r02g.err
timeout 10 python tools/bench_brief.py gpurun_out/r02g_bunnies_q.json
timeout 300 python bench.py --workload cube_drop_4m --quant --no-weak --no-cpu-baseline --no-parity --no-e2e > gpurun_out/r02g_bench4_q.json 2>> gpurun_out/r02g.err
timeout 10 python tools/bench_brief.py gpurun_out/r02g_bench4_q.json
