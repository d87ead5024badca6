# round-2 GPU job 14: control (pf0 vs templated build), pipeline depth sweep, continuous refill reference, bench
mkdir -p gpurun_out
run() { echo -n "$1 games=$2: "; AZB200_LIB=build/variants/lib_$1.so timeout 90 python scripts/profile_selfplay.py $2 | sed 's/.*device_ms.: \([0-9.]*\).*/\1 ms/'; }
for rep in 1 2; do run pf0 4096; run t 4096; done
run t 32768
for d in 1 2 3 4; do timeout 120 python scripts/pipeline_ab.py $d 8 2>&1 | tail -1; done
AZB200_ROUND_TIMES=1 timeout 90 python scripts/bench_configs.py config3 2>&1 | tail -1 | cut -c1-200
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/j14_bench.log 2> gpurun_out/j14_bench.err; tail -c 600 gpurun_out/j14_bench.log; tail -5 gpurun_out/j14_bench.err
