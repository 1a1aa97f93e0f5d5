set -x
mkdir -p gpurun_out
N=$1

timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-weak --no-e2e > gpurun_out/r2p${N}_bench.json 2> gpurun_out/r2p${N}_bench.err
tail -5 gpurun_out/r2p${N}_bench.err
