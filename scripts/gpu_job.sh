# round-2 GPU job 46: round kernel occupancy at 8192 slots; slots beyond the games (no effect expected); round budget re-sweep
mkdir -p gpurun_out
for v in rw28 rw32 rw40 rw28 rw32 rw40; do echo -n "$v "; AZB200_LIB=build/variants/lib_$v.so timeout 120 python scripts/bench_configs.py config3 2>&1 | tail -1 | cut -c80-140; done > gpurun_out/j46_rw.log 2>&1; cat gpurun_out/j46_rw.log
for b in 8 12 16 6 8; do echo -n "round_sims=$b "; AZB200_ROUND_SIMS=$b timeout 120 python scripts/bench_configs.py config3 2>&1 | tail -1 | cut -c80-140; done > gpurun_out/j46_budget.log 2>&1; cat gpurun_out/j46_budget.log
