set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/j1_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/j1_tests.log
AZB200_LIB=build/variants/lib_verify.so python -m pytest tests/test_mcts_gpu.py tests/test_selfplay_gpu.py tests/test_arena_gpu.py -m gpu -x -q > gpurun_out/j1_verify.log 2>&1; echo "verify rc=$?" >> gpurun_out/j1_verify.log
bash scripts/ab.sh run r1 new > gpurun_out/j1_ab.log 2>&1
python scripts/ply_times.py > gpurun_out/j1_ply.log 2>&1
tail -3 gpurun_out/j1_tests.log gpurun_out/j1_verify.log; cat gpurun_out/j1_ab.log
