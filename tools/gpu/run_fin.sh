mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "^Voxelizer\|^$" | tail -30 > gpurun_out/r2fin_tests.log
tail -3 gpurun_out/r2fin_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r2fin_smoke.log 2>&1
tail -2 gpurun_out/r2fin_smoke.log
timeout 900 python bench.py > gpurun_out/r2fin_bench.json 2> gpurun_out/r2fin_bench.err
timeout 10 python tools/bench_brief.py gpurun_out/r2fin_bench.json
