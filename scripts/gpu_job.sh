# round-2 GPU job 49: run-to-run spread of config 3 with the two-tile units, and where it sits
mkdir -p gpurun_out
export AZB200_LIB=build/variants/lib_u9.so
for i in 1 2 3 4 5 6; do AZB200_ROUND_TIMES=1 timeout 120 python scripts/bench_configs.py config3 2>&1 | tail -12 | awk '/slice/{k+=$7*$4; f+=$11*$4; n+=$4} /device_s/{match($0,/"device_s": [0-9.]+/); print substr($0,RSTART,RLENGTH), "k_round_s", k/1e6, "forward_s", f/1e6, "rounds", n}'; done > gpurun_out/j49_spread.log 2>&1
cat gpurun_out/j49_spread.log
nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv
