set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity2.py -m gpu -q -s -k "config2 or config1 or config3" 2>&1 | grep "substep\|passed\|failed\|Error" > gpurun_out/r2b_tests.log
timeout 900 python bench.py --no-weak > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_12m.csv python tools/profile_bench.py > gpurun_out/r2b_prof1.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_p2g3|k_g2p" -s 2 -c 2 -o gpurun_out/r2b_prof python tools/profile_bench.py > gpurun_out/r2b_prof2.log 2>&1
ls -la gpurun_out | tail -5
