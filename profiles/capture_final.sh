#!/bin/bash
# ncu --set full of the main kernels; only compact CSV extracts travel back (reports stay in /tmp on the box).
KEYS='gpu__time_duration.sum|launch__grid_size|launch__block_size|launch__registers_per_thread|launch__shared_mem_per_block_dynamic|launch__occupancy_limit|sm__warps_active.avg.pct_of_peak_sustained_active|smsp__issue_active.avg.pct|smsp__inst_executed.sum|dram__bytes_read.sum |dram__bytes_write.sum |dram__bytes_read.sum$|dram__bytes_write.sum$|gpu__dram_throughput.avg.pct|lts__throughput.avg.pct|lts__t_sector_hit_rate.pct|sm__pipe_tensor_cycles_active.avg.pct|sm__inst_executed_pipe_(xu|fma|alu|fp64|lsu|tmem|tma|uniform).avg.pct_of_peak_sustained_active|sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active|smsp__average_warps_issue_stalled_.*_per_issue_active.ratio|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|sm__throughput.avg.pct'
cap() {  # name regex cmd...
  name=$1; rx=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:$rx -s ${SKIP:-1} -c 1 -o /tmp/$name "$@" > /dev/null 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys,re
rows=list(csv.reader(sys.stdin)); hdr=rows[0]; units=rows[1]; r=rows[2]
rx=re.compile(r'''$KEYS''')
print('metric,unit,value')
print('kernel,,\"%s\"' % r[hdr.index('Kernel Name')])
for k,u,v in zip(hdr,units,r):
    if rx.search(k): print('%s,%s,%s' % (k,u,v))
" > gpurun_out/ncu_final_$name.csv
}
cap k1_smem emcee_smem python profiles/prof_run.py rosenbrock2d 100 0
SKIP=2 cap k3_logistic_tc logistic_tc_kernel python profiles/prof_run.py logistic32d 2 0
KMC_TC=1 cap k2f_gaussian_fused gaussian_fused python profiles/prof_run.py gaussian100d 20 0
cap k1b_bulk emcee_bulk python profiles/prof_run.py gaussian10d 6 0
wc -l gpurun_out/ncu_final_*.csv
