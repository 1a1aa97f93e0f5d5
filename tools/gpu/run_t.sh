mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/r2t_dist.log 2>&1
tail -5 gpurun_out/r2t_dist.log
MPM_COMM=peer MPM_DIST_BACKEND=gloo MPM_SCENE=rebalance timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=3 --master-addr 127.0.0.1 --master-port 29778 tests/dist_worker.py > gpurun_out/r2t_worker.log 2>&1
grep "rebalance\|DIST_\|comm " gpurun_out/r2t_worker.log
