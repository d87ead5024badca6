# round-2 GPU job 56: tower with groups of 4 tiles taken through all layers (depth-first by groups)
mkdir -p gpurun_out
export AZB200_LIB=build/variants/lib_groups.so
timeout 600 python -m pytest tests/test_nnet_gpu.py tests/test_arena_gpu.py tests/test_fullsize_parity_gpu.py -x -q --timeout=300 --timeout-method=thread 2>&1 | tail -3
echo "== groups, tower for every size"; AZB200_TOWER_MAX=100000 timeout 120 python scripts/forward_sweep.py 6 100 | tail -8
echo "== head (layer by layer above 3072)"; AZB200_LIB=build/variants/lib_head2.so timeout 120 python scripts/forward_sweep.py 6 100 | tail -8
for v in head2 groups head2 groups head2 groups; do echo -n "$v "; AZB200_LIB=build/variants/lib_$v.so timeout 120 python scripts/bench_configs.py config3 2>&1 | tail -1 | cut -c100-125; done
