#!/bin/bash
set -u
mkdir -p gpurun_out
for cps in 2 3 4 5; do for ch in 3 4 6 8; do
  echo -n "host path=inc cps=$cps chunks=$ch: "
  PCGRL_INC_CPS=$cps PCGRL_HOST_PATH=inc PCGRL_HOST_CHUNKS=$ch timeout 120 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-configs 2>>gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('e2e %.4g value %.4g' % (d['e2e']['value'], d['value']))"
done; done | tee gpurun_out/r02f_e2e_cps_chunks.txt
tail -3 gpurun_out/ab.err
