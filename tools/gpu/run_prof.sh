set -x
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_12m.csv python tools/profile_bench.py > gpurun_out/r02_prof1.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -s 9 -c 9 -o gpurun_out/r02_prof python tools/profile_bench.py > gpurun_out/r02_prof2.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_g2p2g_12m.csv python tools/profile_bench.py --g2p2g > gpurun_out/r02_prof3.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_p2g3|k_g2p2g_keys" -s 2 -c 2 -o gpurun_out/r02_prof_g2p2g python tools/profile_bench.py --g2p2g > gpurun_out/r02_prof4.log 2>&1
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --workload cube_drop_4m --no-weak > gpurun_out/r02_bench_n1_cfg1.json 2>> gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.json 2>> gpurun_out/r02_bench_n1.err
