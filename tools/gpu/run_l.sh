set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "^Voxelizer\|^$" | tail -30 > gpurun_out/r2l_tests.log
timeout 900 python bench.py --no-weak --no-cpu-baseline --no-e2e > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"k_p2g3|k_g2p" -s 2 -c 2 -o gpurun_out/r02_prof_100m python tools/profile_bench.py --workload multimat_100m > gpurun_out/r2l_prof.log 2>&1
