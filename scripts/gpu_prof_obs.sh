#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_observe_staged -s 3 -c 1 -o gpurun_out/r02_prof_obs -f \
    python scripts/bench_observe.py --reps 2 > gpurun_out/r02_ncu_obs.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02_prof_obs.ncu-rep 2>&1 | head -40
ncu -i gpurun_out/r02_prof_obs.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]
want=['smsp__average_warp_latency_issue_stalled','smsp__average_warps_issue_stalled','l1tex__data_pipe_lsu_wavefronts_mem_shared','smsp__warp_issue_stalled','lts__t_bytes','dram__throughput','l1tex__throughput','sm__warps_active','achieved_occupancy','smsp__pcsamp']
for i,c in enumerate(h):
    if any(w in c for w in want) and 'pct' in c or 'stalled' in c and 'ratio' in c:
        print(c, rows[2][i] if len(rows)>2 else '')
" | sort -t' ' -k2 -g -r | head -40
ncu -i gpurun_out/r02_prof_obs.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_obs.csv 2>/dev/null
python scripts/ncu_lines.py /tmp/src_obs.csv 30 | tee gpurun_out/r02_obs_hot_lines.txt | cut -c1-170
