#!/bin/bash
# ncu --set full of the small / mid-size shard kernels: k_step_lanegroup at 4 Ki envs, k_step_inc at 64 Ki envs
set -u
mkdir -p gpurun_out
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_step_lanegroup -s 20 -c 2 -o gpurun_out/r02_prof_lg -f \
    python bench.py --envs 4096 --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r02_prof_lg.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_step_inc -s 20 -c 2 -o gpurun_out/r02_prof_inc64k -f \
    python bench.py --envs 65536 --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r02_prof_inc64k.log 2>&1
{ echo "# k_step_lanegroup<8>, binary-narrow 16x16, 4 096 envs (ncu --set full --clock-control none, 2 launches)"; python scripts/ncu_summary.py gpurun_out/r02_prof_lg.ncu-rep;
  echo; echo "# k_step_inc<8,1>, binary-narrow 16x16, 65 536 envs"; python scripts/ncu_summary.py gpurun_out/r02_prof_inc64k.ncu-rep; } > gpurun_out/r02_small_shard_ncu_summary.txt 2>&1
cat gpurun_out/r02_small_shard_ncu_summary.txt | cut -c1-130
ncu -i gpurun_out/r02_prof_lg.ncu-rep --page source --csv --print-source cuda,sass -k regex:k_step_lanegroup > /tmp/src_lg.csv 2>/dev/null
python scripts/ncu_lines.py /tmp/src_lg.csv 30 > gpurun_out/r02_lanegroup_hot_lines.txt 2>&1; head -45 gpurun_out/r02_lanegroup_hot_lines.txt | cut -c1-160
rm -f gpurun_out/r02_prof_lg.ncu-rep gpurun_out/r02_prof_inc64k.ncu-rep
