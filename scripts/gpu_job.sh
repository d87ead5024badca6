# round-2 GPU job 65: ncu --set full of one k_tower_tc3 launch in the middle of config 3 (graph off), and of a 1014-position pass
mkdir -p gpurun_out
export AZB200_LIB=build/variants/lib_prof_r2d.so AZB200_GRAPH=0
timeout 800 ncu --set full --clock-control none --import-source on -k regex:k_tower_tc3 --launch-skip 2500 --launch-count 1 -f -o gpurun_out/r2_tower_c3 python scripts/profile_rounds.py 8192 400 > gpurun_out/j65_ncu1.log 2>&1; echo "ncu1 rc=$?"; tail -2 gpurun_out/j65_ncu1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tower_tc3 --launch-skip 5 --launch-count 1 -f -o gpurun_out/r2_tower_1014 python -c "
import importlib,sys; sys.path.insert(0,'.'); azb=importlib.import_module('alphazero-rs_b200'); n=azb.NNet(seed=7,blocks=6); print(n.benchmark(1014,8))" > gpurun_out/j65_ncu2.log 2>&1; echo "ncu2 rc=$?"; tail -2 gpurun_out/j65_ncu2.log
