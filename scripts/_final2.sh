timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 200 2>&1 | tail -3
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -2 gpurun_out/bench_2gpu.err; cut -c1-300 gpurun_out/bench_2gpu.json
