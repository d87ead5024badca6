# round-2 GPU job 24: + one accumulator-release arrival per epilogue warp
mkdir -p gpurun_out
{
for v in mmaw epi mmaw epi; do echo "== $v"; AZB200_LIB=build/variants/lib_$v.so timeout 120 python scripts/forward_sweep.py 6 100; done
echo "== timers (epi), batch 1014 then 8192"
for b in 1014 8192; do AZB200_LIB=build/variants/lib_epi.so AZB200_TC_DEBUG=1 timeout 120 python -c "
import importlib,sys; sys.path.insert(0,'.'); azb=importlib.import_module('alphazero-rs_b200'); n=azb.NNet(seed=7,blocks=6); print(n.benchmark($b,8))" 2>&1 | grep "rank 0\|^[0-9]"; done
} > gpurun_out/j24_sweep.log 2>&1
cat gpurun_out/j24_sweep.log
AZB200_LIB=build/variants/lib_epi.so timeout 600 python -m pytest tests/test_nnet_gpu.py tests/test_train_gpu.py tests/test_train_blocks_gpu.py -x -q --timeout=300 --timeout-method=thread 2>&1 | tail -5
for v in mmaw epi mmaw epi; do echo -n "$v "; AZB200_LIB=build/variants/lib_$v.so timeout 120 python scripts/bench_configs.py config3 2>&1 | tail -1 | cut -c1-200; done
