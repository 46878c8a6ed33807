timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -2 gpurun_out/bench_1gpu.err; cut -c1-160 gpurun_out/bench_1gpu.json
