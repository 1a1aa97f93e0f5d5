mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_dist.py -x -q -m gpu -k "peer_path or empty_rank or two_ranks or rebalance" > gpurun_out/r2xs2_dist.log 2>&1
tail -3 gpurun_out/r2xs2_dist.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2n2_bench.json 2> gpurun_out/r2n2_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2n2_bench.json').read().strip().splitlines()[-1])
w=d['weak_cfg5']
print(round(d['value']/1e9,2), round(d['ms_per_step'],3), 'weak', round(w['value']/1e9,2), round(w['ms_per_step'],3), 'e2e', round(d['e2e']['value']/1e9,2))
for r in d['roofline']['kernel_ms_per_rank']: print(r)
PY
