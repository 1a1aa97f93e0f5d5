set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^Voxelizer\|^$" | tail -60 > gpurun_out/r2c_tests.log
timeout 900 python bench.py --no-weak > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_12m.csv python tools/profile_bench.py > gpurun_out/r2c_prof1.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_p2g3|k_g2p" -s 2 -c 2 -o gpurun_out/r2c_prof python tools/profile_bench.py > gpurun_out/r2c_prof2.log 2>&1
ls -la gpurun_out | tail -5
