set -x
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02f_launches_12m.csv python tools/profile_bench.py > gpurun_out/r02f_prof1.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -s 9 -c 9 -o gpurun_out/r02f_prof python tools/profile_bench.py > gpurun_out/r02f_prof2.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_p2g3|k_g2p" -s 2 -c 2 -o gpurun_out/r02f_prof_100m python tools/profile_bench.py --workload multimat_100m > gpurun_out/r02f_prof3.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_p2g3|k_g2p|k_bin_keys" -s 3 -c 3 -o gpurun_out/r02f_prof_quant python tools/profile_bench.py --workload cube_drop_4m --quant --steps 3 --batch 1 > gpurun_out/r02f_prof4.log 2>&1
tail -2 gpurun_out/r02f_prof*.log
