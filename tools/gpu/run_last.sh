mkdir -p gpurun_out
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
timeout 200 python -m pytest tests/test_gpu_substep.py tests/test_gpu_modes.py -x -q -m gpu 2>&1 | tail -2
timeout 100 python bench.py --workload multimat_12m --no-weak --no-cpu-baseline --no-e2e --repeats 3 > gpurun_out/r2last_12m.json 2> gpurun_out/r2last.err
timeout 10 python tools/bench_brief.py gpurun_out/r2last_12m.json
