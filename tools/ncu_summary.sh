#!/bin/bash
# usage: ncu_summary.sh file.ncu-rep
ncu -i "$1" --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; units=rows[1]
keys=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','launch__shared_mem_per_block_dynamic','launch__grid_size','launch__block_size','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','sm__cycles_elapsed.max','sm__cycles_active.avg']
stall=[h for h in hdr if 'average_warps_issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    for k in keys:
        if k in d: print(f'  {k:75s} {d[k]} {units[hdr.index(k)]}')
    st=sorted(((float(d[k].replace(',','')),k) for k in stall if d.get(k)),reverse=True)
    for v,k in st[:9]: print(f'  {k:95s} {v:.3f}')
"
