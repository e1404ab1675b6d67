#!/bin/bash
# time the search-kernel variants under gpurun_variants/ on the smb / sokoban workloads
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in gpurun_variants/lib*.so; do
  name=$(basename $lib .so)
  for wl in smb-narrow-116x16 sokoban-cellular-5x5 sokoban-narrow-5x5; do
    PCGRL_B200_LIB=$PWD/$lib timeout 150 python bench.py --workload $wl --steps ${STEPS:-40} --warmup 4 --no-cpu-baseline --no-e2e 2>>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', '$wl', 'value %.4g'%d['value'], 'ms/step %.3f'%d['ms_per_step'])"
  done
done | tee gpurun_out/ab_search.txt
