# round-2 GPU job 42: stem with fewer working CTAs at round sizes
mkdir -p gpurun_out
AZB200_LIB=build/variants/lib_stem.so timeout 400 python -m pytest tests/test_nnet_gpu.py tests/test_train_gpu.py -x -q --timeout=200 --timeout-method=thread 2>&1 | tail -3
for v in head stem head stem; do echo "== $v"; AZB200_LIB=build/variants/lib_$v.so timeout 120 python scripts/forward_sweep.py 6 100 | tail -8; done > gpurun_out/j42_sweep.log 2>&1; head -18 gpurun_out/j42_sweep.log
for v in head stem head stem; do echo -n "$v "; AZB200_LIB=build/variants/lib_$v.so timeout 120 python scripts/bench_configs.py config3 2>&1 | tail -1 | cut -c80-200; done
