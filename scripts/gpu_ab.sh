#!/bin/bash
set -u
mkdir -p gpurun_out
for dc in 0 2 3 4 6; do for cps in 4 8; do
PCGRL_DEVICE_CHUNKS=$dc PCGRL_INC_CPS=$cps timeout 200 python bench.py --steps 400 --warmup 10 --no-cpu-baseline --no-configs --no-e2e 2>>gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('device chunks $dc cps $cps: value %.4g ms/step %.4f' % (d['value'], d['ms_per_step']))"
done; done
tail -3 gpurun_out/ab.err
