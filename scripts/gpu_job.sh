# round-2 GPU job 36: tower on position-aligned tiles, pair-local ordering only
mkdir -p gpurun_out
export AZB200_LIB=build/variants/lib_tower8.so
timeout 300 python -m pytest tests/test_nnet_gpu.py -x -q --timeout=120 --timeout-method=thread 2>&1 | tail -5
for b in 1014 338; do AZB200_TOWER_DEBUG=1 timeout 120 python -c "
import importlib,sys; sys.path.insert(0,'.'); azb=importlib.import_module('alphazero-rs_b200'); n=azb.NNet(seed=7,blocks=6); print(n.benchmark($b,8))" 2>&1 | grep "tower\|^[0-9]"; done > gpurun_out/j36_timeline.log 2>&1
cat gpurun_out/j36_timeline.log
timeout 120 python scripts/forward_sweep.py 6 100 2>&1 | tail -8 > gpurun_out/j36_sweep_tower.log; cat gpurun_out/j36_sweep_tower.log
AZB200_TOWER=0 timeout 120 python scripts/forward_sweep.py 6 100 2>&1 | tail -8 > gpurun_out/j36_sweep_layers.log; cat gpurun_out/j36_sweep_layers.log
for t in 1 0 1 0; do echo -n "tower=$t "; AZB200_TOWER=$t AZB200_TIMING=1 timeout 120 python scripts/bench_configs.py config3 2>&1 | grep -i "capture failed\|device_s" | cut -c1-200; done
