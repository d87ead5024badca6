python -m pytest tests/test_nnet_gpu.py tests/test_arena_gpu.py tests/test_learn_gpu.py -x -q -m gpu 2>&1 | tail -4
for d in 1 0; do AZB200_LEAF_DEDUP=$d python scripts/profile_nn_selfplay.py 8192 400 6 2>&1 | tail -2; done
