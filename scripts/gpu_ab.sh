#!/bin/bash
set -u
mkdir -p gpurun_out
for chunks in 4 6; do for taper in 0 40 70; do for rep in 1 2; do
  PCGRL_HOST_CHUNKS=$chunks PCGRL_HOST_TAPER=$taper timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-configs 2>>gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunks $chunks taper $taper: e2e %.4g value %.4g' % (d['e2e']['value'], d['value']))"
done; done; done
tail -3 gpurun_out/ab.err
