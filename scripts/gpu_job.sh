# round-2 GPU job 50: final validation on the final tree: GPU suite, smoke, bench (own arm + reference arm), config-3 spread
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout=400 --timeout-method=thread > gpurun_out/j50_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/j50_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/j50_ref.log 2>&1; echo "ref rc=$?"; cut -c1-300 gpurun_out/j50_ref.log | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/j50_bench.log 2> gpurun_out/j50_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/j50_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/j50_bench.log') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['clocks'])
print('nnet_forward', d['nnet_forward'].get('ms_per_pass'), d['nnet_forward'].get('roofline',{}).get('frac'))
c3=d['config3']; print('config3', c3.get('device_s'), c3.get('roofline',{}).get('frac'), c3.get('e2e',{}).get('value'), c3.get('parity_checked'))
print('config4', d['config4'].get('device_s_max_over_ranks'), 'config5', {k:d['config5'].get(k) for k in ('wall_s_rank0','selfplay_s','train_s','arena_s')})
print('cpu', d.get('cpu_baseline',{}).get('value'), 'ratio e2e/cpu', d['e2e']['value']/d['cpu_baseline']['value'])
PY
for i in 1 2 3 4 5 6 7 8; do timeout 120 python scripts/bench_configs.py config3 2>&1 | tail -1 | cut -c100-130; done > gpurun_out/j50_c3_spread.log; tr '\n' ' ' < gpurun_out/j50_c3_spread.log
