mkdir -p gpurun_out
for g in 6 2 3 1 7; do
MPM_G2P_CFG=$g timeout 600 python bench.py --workload multimat_12m --no-weak --no-cpu-baseline --no-parity --no-e2e --repeats 4 > gpurun_out/r2i_g2p$g.json 2>> gpurun_out/r2i.err
done
for p in 1 2 3 4; do
MPM_P2G_CFG=$p timeout 600 python bench.py --workload multimat_12m --no-weak --no-cpu-baseline --no-parity --no-e2e --repeats 4 > gpurun_out/r2i_p2g$p.json 2>> gpurun_out/r2i.err
done
