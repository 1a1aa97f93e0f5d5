mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 500 $CS --tool memcheck --print-limit 20 python tools/sanitize_scene.py > gpurun_out/r2w_memcheck.log 2>&1
tail -4 gpurun_out/r2w_memcheck.log
timeout 700 $CS --tool racecheck --print-limit 20 python tools/sanitize_scene.py > gpurun_out/r2w_racecheck.log 2>&1
tail -4 gpurun_out/r2w_racecheck.log
MPM_G2P_TILE=1 timeout 300 $CS --tool memcheck --print-limit 20 python tools/sanitize_scene.py > gpurun_out/r2w_memcheck_bulk.log 2>&1
tail -3 gpurun_out/r2w_memcheck_bulk.log
