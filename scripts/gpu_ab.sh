#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "maze3d or holey or trace or fixtures or search" 2>&1 | tail -3
for r in 1 2; do
timeout 300 python bench.py --workload minecraft_3D_maze-narrow-14x14x14 --steps 40 --warmup 4 --no-cpu-baseline --no-configs --no-e2e 2>>gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('maze3d: value %.4g kernel_ms %.3f' % (d['value'], d['roofline']['kernel_ms_per_launch']))"
done
tail -3 gpurun_out/ab.err
