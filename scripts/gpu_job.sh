# round-2 GPU job 68: launch list of the bench command on the final build (ncu --metrics gpu__time_duration.sum, first 400 launches)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/j68_b.log 2>&1; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/j68_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/j68_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/j68_launches.csv')) if len(r)>10]
h=rows[0]; k=h.index('Kernel Name'); v=h.index('Metric Value')
t=collections.Counter(); n=collections.Counter()
for r in rows[1:]:
    name=r[k].split('(')[0][:40]; t[name]+=float(r[v].replace(',','')); n[name]+=1
tot=sum(t.values())
for name,x in t.most_common(12): print(f"{name:42s} x{n[name]:4d} {x/1e3:10.1f} us {100*x/tot:5.1f}%")
PY
