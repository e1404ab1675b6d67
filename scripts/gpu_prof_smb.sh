#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_smb_solve -s 5 -c 1 -o gpurun_out/r02_prof_smb -f \
    python bench.py --workload smb-narrow-116x16 --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r02_ncu_smb.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02_prof_smb.ncu-rep 2>&1 | head -40
ncu -i gpurun_out/r02_prof_smb.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_smb.csv 2>/dev/null
python scripts/ncu_lines.py /tmp/src_smb.csv 45 | tee gpurun_out/r02_smb_hot_lines.txt | cut -c1-180
