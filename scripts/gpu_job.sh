# round-2 GPU job 57: stem launched with programmatic serialization behind k_round (table copy under the round kernel's tail)
mkdir -p gpurun_out
export AZB200_LIB=build/variants/lib_stempdl.so
timeout 600 python -m pytest tests/test_nnet_gpu.py tests/test_arena_gpu.py tests/test_rounds_stress_gpu.py -x -q --timeout=300 --timeout-method=thread 2>&1 | tail -3
for v in head3 stempdl head3 stempdl head3 stempdl; do echo -n "$v "; AZB200_LIB=build/variants/lib_$v.so timeout 200 python scripts/bench_configs.py config3 config4 2>&1 | tail -2 | cut -c100-135 | tr '\n' ' '; echo; done
