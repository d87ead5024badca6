# round-2 GPU job 21: resident warps per SM of the round kernel (config 3), A/B on one box
mkdir -p gpurun_out
for rep in 1 2; do for w in 28 32 36 40; do
  echo -n "rw$w "; AZB200_LIB=build/variants/lib_rw$w.so timeout 120 python scripts/bench_configs.py config3 2>&1 | tail -1 | cut -c1-330
done; done > gpurun_out/j21_rw.log 2>&1
cat gpurun_out/j21_rw.log
