# round-2 GPU job 67: stress of the final build: config-3 sized calls (6 blocks, 8192 games x 100 sims) repeated, dedup/cache on vs off,
# tower on vs off: every run must produce the same games
mkdir -p gpurun_out
{
timeout 500 python scripts/stress_dedup.py 8192 100 6 5
echo "--- tower off"; AZB200_TOWER=0 timeout 300 python scripts/stress_dedup.py 8192 100 6 1
echo "--- 4736 slots"; AZB200_ROUND_SLOTS=4736 timeout 300 python scripts/stress_dedup.py 8192 100 6 1
} > gpurun_out/j67_stress.log 2>&1
cat gpurun_out/j67_stress.log | cut -c1-200
