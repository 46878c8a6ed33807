set -x
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -2 gpurun_out/bench_1gpu.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -2 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-agent > gpurun_out/b_ncu.log 2>&1; tail -1 gpurun_out/b_ncu.log | cut -c1-200
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:rollout_kernel_persist -s 1 -c 1 --csv --log-file gpurun_out/traffic.csv python scripts/profile_rollout.py 512 128 2001 2 > gpurun_out/traffic.log 2>&1; tail -4 gpurun_out/traffic.csv
