# round-2 GPU job 27: weight / activation fetch order (first 6 k-blocks, A tile, the rest)
mkdir -p gpurun_out
AZB200_LIB=build/variants/lib_wfirst.so timeout 600 python -m pytest tests/test_nnet_gpu.py tests/test_train_gpu.py -x -q --timeout=300 --timeout-method=thread 2>&1 | tail -3
{
for v in epi wfirst epi wfirst; do echo "== $v"; AZB200_LIB=build/variants/lib_$v.so timeout 120 python scripts/forward_sweep.py 6 100; done
echo "== timers (wfirst), batch 1014"
AZB200_LIB=build/variants/lib_wfirst.so AZB200_TC_DEBUG=1 timeout 120 python -c "
import importlib,sys; sys.path.insert(0,'.'); azb=importlib.import_module('alphazero-rs_b200'); n=azb.NNet(seed=7,blocks=6); print(n.benchmark(1014,8))" 2>&1 | grep "rank 0\|^[0-9]"
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
} > gpurun_out/j27_wfirst.log 2>&1
cat gpurun_out/j27_wfirst.log
