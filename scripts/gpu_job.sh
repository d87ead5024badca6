# round-2 GPU job 58: validation of the final tree: GPU suite, smoke, bench (own arm + reference arm)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout=400 --timeout-method=thread > gpurun_out/j58_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/j58_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py --impl reference > gpurun_out/j58_ref.log 2>&1; echo "ref rc=$?"; cut -c1-200 gpurun_out/j58_ref.log | tail -1
timeout 600 python bench.py > gpurun_out/j58_bench.log 2> gpurun_out/j58_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/j58_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/j58_bench.log') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','steps','warmup')}, d['e2e']['value'], d['roofline']['frac'], d['clocks'])
c3=d['config3']; print('config3', c3.get('device_s'), 'config4', d['config4'].get('device_s_max_over_ranks'), 'config5', d['config5'].get('wall_s_rank0'))
print('cpu', d.get('cpu_baseline',{}).get('value'), 'ratio e2e/cpu', d['e2e']['value']/d['cpu_baseline']['value'])
PY
