# round-2 GPU job 53: tower with equal position shares per pair (two-tile units + one single tile)
mkdir -p gpurun_out
export AZB200_LIB=build/variants/lib_share.so
timeout 600 python -m pytest tests/test_nnet_gpu.py tests/test_arena_gpu.py tests/test_fullsize_parity_gpu.py -x -q --timeout=300 --timeout-method=thread 2>&1 | tail -3
for v in pair share pair share; do echo "== $v"; AZB200_LIB=build/variants/lib_$v.so timeout 120 python scripts/forward_sweep.py 6 100 | tail -8; done > gpurun_out/j53_sweep.log 2>&1; head -18 gpurun_out/j53_sweep.log
for v in pair share pair share pair share; do echo -n "$v "; AZB200_LIB=build/variants/lib_$v.so timeout 120 python scripts/bench_configs.py config3 config4 2>&1 | tail -2 | cut -c1-40,100-125 | tr '\n' ' '; echo; done
