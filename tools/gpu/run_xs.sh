mkdir -p gpurun_out
for v in xs base xs base; do
L=$PWD/taichi_elements_b200/libmpm_b200_$v.so; [ $v = base ] && L=$PWD/taichi_elements_b200/libmpm_b200.so
MPM_B200_LIB=$L timeout 200 python bench.py --workload multimat_12m --no-weak --no-cpu-baseline --no-e2e --repeats 4 > gpurun_out/r2xs_$v.json 2>> gpurun_out/r2xs.err
echo $v; timeout 10 python tools/bench_brief.py gpurun_out/r2xs_$v.json; python -c "
import json; d=json.loads(open('gpurun_out/r2xs_$v.json').read().strip().splitlines()[-1]); print(d['parity_check']['one_substep_rel_err'])"
done
tail -3 gpurun_out/r2xs.err
