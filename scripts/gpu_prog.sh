#!/bin/bash
# GPU-box run for the progressive host pipeline: parity of the new path, its device timeline, and the e2e arm of
# bench.py against the chunk-per-stream pipeline for several chunk counts / search CTAs per SM.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_split.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or compact" 2>&1 | tail -5
PCGRL_HOST_TRACE=2 timeout 300 python scripts/host_trace.py 2>&1 | grep -v "^call 3" | head -40 | tee gpurun_out/prog_host_trace.txt
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-configs > gpurun_out/prog_$tag.json 2> gpurun_out/prog_$tag.err
  python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/prog_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "device %.3e" % l["value"], "e2e %.3e" % l["e2e"]["value"], "launches", l.get("gpu_launches"))
except Exception as exc:
    print("$tag", "FAILED", exc)
PY
}
run old PCGRL_HOST_PROG=0
run prog_c8 PCGRL_HOST_PROG=1
run prog_c4 PCGRL_HOST_PROG=1 PCGRL_HOST_CHUNKS=4
run prog_c16 PCGRL_HOST_PROG=1 PCGRL_HOST_CHUNKS=16
run prog_c8_cps5 PCGRL_HOST_PROG=1 PCGRL_PROG_CPS=5
run prog_c8_cps7 PCGRL_HOST_PROG=1 PCGRL_PROG_CPS=7
run prog_c12 PCGRL_HOST_PROG=1 PCGRL_HOST_CHUNKS=12
tail -3 gpurun_out/prog_prog_c8.err
# lane-group step against the other one-launch / three-launch paths by shard size
sweep() {
  n=$1; path=$2; shift 2
  env PCGRL_STEP_PATH=$path "$@" timeout 200 python bench.py --envs $n --steps 300 --warmup 10 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/lg_tmp.json 2> gpurun_out/lg_tmp.err
  python - "$n" "$path" "$*" <<'PY'
import json, sys
try:
    l = json.loads(open("gpurun_out/lg_tmp.json").read().strip().splitlines()[-1])
    print("envs=%s path=%s %s: value %.4g ms_per_step %.4f kernel %s" % (sys.argv[1], sys.argv[2], sys.argv[3], l["value"], l["ms_per_step"], l["roofline"]["kernel"]))
except Exception as exc:
    print("envs=%s path=%s FAILED %s" % (sys.argv[1], sys.argv[2], exc)); print(open("gpurun_out/lg_tmp.err").read()[-600:])
PY
}
{
sweep 16384 incfused
sweep 16384 lg
sweep 65536 incfused
sweep 65536 lg
sweep 65536 lg PCGRL_LG_TILE=16
sweep 65536 lg PCGRL_LG_TILE=32
sweep 131072 incfused
sweep 131072 lg
sweep 262144 inc
sweep 262144 lg
sweep 524288 inc
sweep 524288 lg
} | tee gpurun_out/lg_paths_by_size.txt
