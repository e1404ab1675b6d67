#!/bin/bash
set -u
mkdir -p gpurun_out
for last in 4 6 8; do for r in 1 2; do
PCGRL_INC_CPS_LAST=$last timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-configs 2>>gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('last-chunk cps $last: e2e %.4g value %.4g' % (d['e2e']['value'], d['value']))"
done; done
PCGRL_HOST_CHUNKS=7 timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-configs 2>>gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunks 7 (last alone, cps 8): e2e %.4g value %.4g' % (d['e2e']['value'], d['value']))"
timeout 900 python -m pytest tests -m gpu -x -q -k "host or packed or compact or pipelined or split or vector" 2>&1 | tail -3
tail -3 gpurun_out/ab.err
