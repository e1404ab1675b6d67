#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_step_search -s 6 -c 1 -o gpurun_out/r02h_prof_maze3d -f \
    python bench.py --workload minecraft_3D_maze-narrow-14x14x14 --steps 10 --warmup 4 --no-e2e --no-cpu-baseline > gpurun_out/r02h_ncu_maze.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02h_prof_maze3d.ncu-rep 2>&1 | head -30
ncu -i gpurun_out/r02h_prof_maze3d.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_maze.csv 2>/dev/null
python scripts/ncu_lines.py /tmp/src_maze.csv 40 | tee gpurun_out/r02h_maze3d_hot_lines.txt | cut -c1-180
