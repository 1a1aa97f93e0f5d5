set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2a_tests.log
timeout 600 python bench.py --no-weak > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
timeout 300 python bench.py --workload multimat_12m --no-weak --no-cpu-baseline --no-parity --no-e2e > gpurun_out/r2a_bench12.json 2>> gpurun_out/r2a_bench.err
timeout 300 python bench.py --workload cube_drop_4m --no-weak --no-cpu-baseline --no-parity > gpurun_out/r2a_bench4.json 2>> gpurun_out/r2a_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 940 -c 40 --csv --log-file gpurun_out/r2a_launches_12m.csv python tools/profile_bench.py > gpurun_out/r2a_prof1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_p2g3|k_g2p" -s 206 -c 2 -o gpurun_out/r2a_prof python tools/profile_bench.py > gpurun_out/r2a_prof2.log 2>&1
ls -la gpurun_out | tail
