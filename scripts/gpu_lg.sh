#!/bin/bash
# GPU-box run for the lane-group step (step_lanegroup.cu): parity of every step path, then ms per step by shard size
# against the one-launch / three-launch thread-per-grid paths.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_split.py -m gpu -x -q 2>&1 | tail -8
sweep() {
  n=$1; path=$2; shift 2
  env PCGRL_STEP_PATH=$path "$@" timeout 200 python bench.py --envs $n --steps 300 --warmup 10 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/lg_tmp.json 2> gpurun_out/lg_tmp.err
  python - "$n" "$path" "$*" <<'PY'
import json, sys
try:
    l = json.loads(open("gpurun_out/lg_tmp.json").read().strip().splitlines()[-1])
    print("envs=%s path=%s %s: value %.4g ms_per_step %.4f" % (sys.argv[1], sys.argv[2], sys.argv[3], l["value"], l["ms_per_step"]))
except Exception as exc:
    print("envs=%s path=%s FAILED %s" % (sys.argv[1], sys.argv[2], exc)); print(open("gpurun_out/lg_tmp.err").read()[-600:])
PY
}
{
sweep 4096 incfused
sweep 4096 lg
sweep 16384 incfused
sweep 16384 lg
sweep 65536 incfused
sweep 65536 lg
sweep 65536 lg PCGRL_LG_TILE=8
sweep 65536 lg PCGRL_LG_TILE=32
sweep 131072 incfused
sweep 131072 lg
sweep 262144 inc
sweep 262144 lg
} | tee gpurun_out/lg_paths_by_size.txt
