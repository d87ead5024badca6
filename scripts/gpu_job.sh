# round-2 GPU job 40: durations of the three kernels of a forward pass (stem, tower, heads) at 64 / 1014 / 3000 positions
mkdir -p gpurun_out
for b in 64 1014 3000; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/j40_fwd_$b.csv python -c "
import importlib,sys; sys.path.insert(0,'.'); azb=importlib.import_module('alphazero-rs_b200'); n=azb.NNet(seed=7,blocks=6); print(n.benchmark($b,4))" > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/j40_fwd_$b.csv')) if len(r)>10]
h=rows[0]; k=h.index('Kernel Name'); v=h.index('Metric Value')
last=rows[-12:]
print('batch $b:', [(r[k][:24], r[v]) for r in rows[-9:]])
PY
done
