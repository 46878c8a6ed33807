timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:rollout_kernel_persist -s 1 -c 1 --csv --log-file gpurun_out/traffic.csv python scripts/profile_rollout.py 512 128 2001 2 > gpurun_out/traffic.log 2>&1; tail -5 gpurun_out/traffic.csv | cut -d, -f13-15
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -2 gpurun_out/bench_1gpu.err; cut -c1-200 gpurun_out/bench_1gpu.json
