#!/bin/bash
# quick counters for the self-play kernel (few passes): bash scripts/ncu_quick.sh <tag> [games] [sims]
TAG=$1; shift
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__cycles_active.avg,sm__cycles_elapsed.avg
mkdir -p gpurun_out
python scripts/profile_selfplay.py "$@" > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:k_selfplay -c 1 --csv --log-file gpurun_out/quick_$TAG.csv python scripts/profile_selfplay.py "$@" > /dev/null 2>&1
cat gpurun_out/plain_$TAG.log
grep -v "^==" gpurun_out/quick_$TAG.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin): print(r['Metric Name'], r['Metric Unit'], r['Metric Value'])
"
