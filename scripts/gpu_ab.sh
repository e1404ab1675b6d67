#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "smb or search or fixtures" 2>&1 | tail -3
for n in 65536 262144; do
timeout 300 python bench.py --workload smb-narrow-116x16 --envs $n --steps 20 --warmup 3 --no-cpu-baseline --no-configs --no-e2e 2>>gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('smb $n envs: value %.4g kernel_ms %.3f' % (d['value'], d['roofline']['kernel_ms_per_launch']))"
done
tail -3 gpurun_out/ab.err
