# round-2 GPU job 54: validation of the tree: GPU suite, smoke, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout=400 --timeout-method=thread > gpurun_out/j54_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/j54_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/j54_bench.log 2> gpurun_out/j54_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/j54_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/j54_bench.log') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['clocks'])
print('nnet_forward', d['nnet_forward'].get('ms_per_pass'), d['nnet_forward'].get('roofline',{}).get('frac'))
c3=d['config3']; print('config3', c3.get('device_s'), c3.get('roofline',{}).get('frac'), c3.get('e2e',{}).get('value'), c3.get('parity_checked'), c3.get('kernel_launches'))
print('config4', d['config4'].get('device_s_max_over_ranks'), 'config5', {k:d['config5'].get(k) for k in ('wall_s_rank0','selfplay_s','train_s','arena_s')})
print('cpu', d.get('cpu_baseline',{}).get('value'), 'ratio e2e/cpu', d['e2e']['value']/d['cpu_baseline']['value'])
PY
