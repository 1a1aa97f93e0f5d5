mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/r2y_dist.log 2>&1
tail -3 gpurun_out/r2y_dist.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-weak > gpurun_out/r2y_n2.json 2> gpurun_out/r2y_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2y_n2.json').read().strip().splitlines()[-1])
print(d['value']/1e9, d['ms_per_step'], d['e2e'])
PY
