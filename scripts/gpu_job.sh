# round-2 GPU job 60: 2 warps per CTA everywhere: GPU suite, config 3 / config 4 against 4 warps per CTA
mkdir -p gpurun_out
AZB200_LIB=build/variants/lib_w2.so timeout 900 python -m pytest tests -x -q -m gpu --timeout=400 --timeout-method=thread 2>&1 | tail -3
for v in w4 w2 w4 w2 w4 w2; do echo -n "$v "; AZB200_LIB=build/variants/lib_$v.so timeout 200 python scripts/bench_configs.py config3 config4 2>&1 | tail -2 | cut -c100-135 | tr '\n' ' '; echo; done
