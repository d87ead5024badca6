# round-2 GPU job 16: full GPU suite + bench (3 deep, 32 warps per SM) + smoke
mkdir -p gpurun_out
T="--timeout=300 --timeout-method=thread"
timeout 700 python -m pytest tests -m gpu -q $T > gpurun_out/j16_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/j16_tests.log
tail -5 gpurun_out/j16_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/j16_bench.log 2> gpurun_out/j16_bench.err; tail -c 400 gpurun_out/j16_bench.log; tail -5 gpurun_out/j16_bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -c 500
