# round-2 GPU job 62: the bench on the final build (2 warps per CTA), own arm
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/j62_bench.log 2> gpurun_out/j62_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/j62_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/j62_bench.log') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','steps','warmup')}, d['e2e']['value'], d['roofline']['frac'], d['clocks'], d['config']['single_batch_ms'])
c3=d['config3']; print('config3', c3.get('device_s'), 'config4', d['config4'].get('device_s_max_over_ranks'), 'config5', d['config5'].get('wall_s_rank0'), 'config1', d['config1']['uniform_50_sims']['device_sims_per_sec'])
print('cpu', d.get('cpu_baseline',{}).get('value'), 'ratio e2e/cpu', d['e2e']['value']/d['cpu_baseline']['value'])
PY
