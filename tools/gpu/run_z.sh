mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist.py -x -q -m gpu -k "two_ranks" > gpurun_out/r2z_dist.log 2>&1
tail -15 gpurun_out/r2z_dist.log
