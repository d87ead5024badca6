# A/B of the leaf de-duplication and the evaluation cache on BASELINE config 3 (identical games in all three runs)
python -m pytest tests/test_nnet_gpu.py tests/test_arena_gpu.py tests/test_learn_gpu.py tests/test_selfplay_gpu.py -x -q -m gpu 2>&1 | tail -4
for v in "1 1" "1 0" "0 0"; do set -- $v; echo "dedup=$1 cache=$2"; AZB200_LEAF_DEDUP=$1 AZB200_EVAL_CACHE=$2 python scripts/profile_nn_selfplay.py 8192 400 6 2>&1 | tail -2; done
