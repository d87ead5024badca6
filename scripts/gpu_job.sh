# round-2 GPU job 10: config 3 in wave mode (K = 2, 4, 8), config 2 tail with waves
mkdir -p gpurun_out
for K in 1 2 4 8; do AZB200_BENCH_THREADS=$K AZB200_ROUND_TIMES=1 timeout 120 python scripts/bench_configs.py config3 2>&1 | tail -12 | cut -c1-420; done > gpurun_out/j10_c3_waves.log 2>&1
cat gpurun_out/j10_c3_waves.log
