# round-2 GPU job 51: the arena's two towers in one launch
mkdir -p gpurun_out
export AZB200_LIB=build/variants/lib_pair.so
timeout 600 python -m pytest tests/test_nnet_gpu.py tests/test_arena_gpu.py tests/test_fullsize_parity_gpu.py tests/test_learn_gpu.py -x -q --timeout=300 --timeout-method=thread 2>&1 | tail -3
for t in 1 0 1 0; do echo -n "pair=$t "; AZB200_TOWER_PAIR=$t timeout 200 python scripts/bench_configs.py config4 2>&1 | tail -1 | cut -c1-260; done
