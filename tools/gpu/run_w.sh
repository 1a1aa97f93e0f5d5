mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 35 $CS --tool memcheck --print-limit 20 python tools/sanitize_scene.py > gpurun_out/r2w_memcheck.log 2>&1
tail -2 gpurun_out/r2w_memcheck.log
timeout 40 $CS --tool racecheck --print-limit 20 python tools/sanitize_scene.py > gpurun_out/r2w_racecheck.log 2>&1
tail -2 gpurun_out/r2w_racecheck.log
