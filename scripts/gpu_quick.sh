#!/bin/bash
# quick GPU check: the whole parity suite, then the 3D maze bench line (its shared memory per warp changed)
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_quick.txt
timeout 300 python bench.py --workload minecraft_3D_maze-narrow-14x14x14 --steps 40 --warmup 4 --no-cpu-baseline 2>>gpurun_out/q.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('maze3d value %.4g e2e %.4g kernel_ms %.3f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_launch']))"
tail -3 gpurun_out/q.err
