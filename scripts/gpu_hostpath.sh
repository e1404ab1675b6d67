#!/bin/bash
# e2e of the headline workload with the host pipeline's chunks on the one-launch incremental kernel instead of the
# three-launch path
set -u
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', 'device %.3e'%l['value'], 'e2e %.3e'%l['e2e']['value'], 'launches', l.get('gpu_launches'))"
}
{
run inc_default
run incfused_cps4 PCGRL_HOST_PATH=incfused
run incfused_cps6 PCGRL_HOST_PATH=incfused PCGRL_INC_CPS=6
run incfused_cps8 PCGRL_HOST_PATH=incfused PCGRL_INC_CPS=8
run incfused_cps6_c3 PCGRL_HOST_PATH=incfused PCGRL_INC_CPS=6 PCGRL_HOST_CHUNKS=3
run inc_default_again
} | tee gpurun_out/e2e_hostpath.txt
