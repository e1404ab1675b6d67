#!/bin/bash
set -u
mkdir -p gpurun_out
for cps in 3 4; do for ch in 5 6 8 10 12 16; do
  echo -n "graph=1 cps=$cps chunks=$ch: "
  PCGRL_INC_CPS=$cps PCGRL_HOST_GRAPH=1 PCGRL_HOST_CHUNKS=$ch timeout 120 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-configs 2>>gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('e2e %.4g value %.4g' % (d['e2e']['value'], d['value']))"
done; done | tee gpurun_out/r02i_e2e_graph_sweep.txt
tail -3 gpurun_out/ab.err
