set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2n2_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2n2_bench.json 2> gpurun_out/r2n2_bench.err
tail -5 gpurun_out/r2n2_bench.err
