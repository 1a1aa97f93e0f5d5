set -x
mkdir -p gpurun_out
for d in 1 0; do
MPM_DEFER_SVD=$d timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_p2g3" -s 2 -c 1 -o gpurun_out/r2f_p2g_d$d python tools/profile_bench.py --workload cube_drop_4m --preroll 10 > gpurun_out/r2f_prof_d$d.log 2>&1
done
