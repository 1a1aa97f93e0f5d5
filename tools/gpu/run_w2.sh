mkdir -p gpurun_out
for v in w2 w3 w4 base w2 w3 base; do
L=$PWD/taichi_elements_b200/libmpm_b200_$v.so; [ $v = base ] && L=$PWD/taichi_elements_b200/libmpm_b200.so
MPM_B200_LIB=$L timeout 200 python bench.py --workload multimat_12m --no-weak --no-cpu-baseline --no-parity --no-e2e --repeats 4 > gpurun_out/r2w2_$v.json 2>> gpurun_out/r2w2.err
echo $v; timeout 10 python tools/bench_brief.py gpurun_out/r2w2_$v.json
done
