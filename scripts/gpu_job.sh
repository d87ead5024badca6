# round-2 GPU job: every stage bounded by its own timeout
mkdir -p gpurun_out
T="--timeout=240 --timeout-method=thread"
timeout 170 python -m pytest tests/test_mcts_gpu.py tests/test_selfplay_gpu.py tests/test_arena_gpu.py -m gpu -x -q $T > gpurun_out/j2_core.log 2>&1; echo "core rc=$?" >> gpurun_out/j2_core.log
tail -3 gpurun_out/j2_core.log
AZB200_LIB=build/variants/lib_verify.so timeout 200 python -m pytest tests/test_mcts_gpu.py tests/test_selfplay_gpu.py tests/test_arena_gpu.py -m gpu -x -q $T > gpurun_out/j2_verify.log 2>&1; echo "verify rc=$?" >> gpurun_out/j2_verify.log
tail -3 gpurun_out/j2_verify.log
timeout 120 bash scripts/ab.sh run r1 new > gpurun_out/j2_ab.log 2>&1; cat gpurun_out/j2_ab.log
timeout 60 python scripts/ply_times.py > gpurun_out/j2_ply.log 2>&1
timeout 420 python -m pytest tests -m gpu -q $T --deselect tests/test_mcts_gpu.py --deselect tests/test_selfplay_gpu.py --deselect tests/test_arena_gpu.py > gpurun_out/j2_rest.log 2>&1; echo "rest rc=$?" >> gpurun_out/j2_rest.log
tail -15 gpurun_out/j2_rest.log
timeout 60 python scripts/bench_configs.py config3 > gpurun_out/j2_c3_graph.log 2>&1; cat gpurun_out/j2_c3_graph.log
AZB200_GRAPH=0 timeout 60 python scripts/bench_configs.py config3 > gpurun_out/j2_c3_nograph.log 2>&1; cat gpurun_out/j2_c3_nograph.log
