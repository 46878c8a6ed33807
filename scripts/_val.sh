export SERL_B200_LIB=$PWD/serl_b200/libserl_b200_v4.so
timeout 400 python -m pytest tests/test_parity_strict_gpu.py tests/test_rollout_gpu.py tests/test_edge_gpu.py tests/test_boundary_gpu.py -m gpu -q -x --timeout 150 2>&1 | tail -5
for m in nominal mixed; do timeout 120 python scripts/profile_rollout.py 512 128 2001 3 $m 2>&1 | tail -1; done
timeout 60 python scripts/profile_rollout.py 64 128 2001 3 2>&1 | tail -1
