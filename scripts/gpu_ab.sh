#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "multiagent" 2>&1 | tail -4
for lib in gpurun_variants/lib*.so; do
  name=$(basename $lib .so); name=${name#lib}
  for rep in 1 2; do
  PCGRL_B200_LIB=$PWD/$lib timeout 200 python bench.py --workload minecraft_3D_maze-narrow-14x14x14 --steps 40 --warmup 4 --no-cpu-baseline --no-e2e --no-configs 2>>gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('maze3d $name: value %.4g kernel_ms %.3f' % (d['value'], d['roofline']['kernel_ms_per_launch']))"
  done
done
tail -3 gpurun_out/ab.err
