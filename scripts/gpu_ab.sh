#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "observ or trace or vector or holey or multiagent or static or wrapped" 2>&1 | tail -3
timeout 300 python scripts/bench_observe.py 2>>gpurun_out/ab.err > gpurun_out/r02_observe.jsonl; python - <<PY
import json
for l in open("gpurun_out/r02_observe.jsonl"):
    d=json.loads(l); print("  %-34s %-16s ms %.4f frac %.3f" % (d["case"], d["dtype"], d["ms"], d["frac_of_hbm_peak"]))
PY
timeout 300 python scripts/bench_rl_loop.py 2>>gpurun_out/ab.err > gpurun_out/r02_rl_loop.jsonl; cat gpurun_out/r02_rl_loop.jsonl
tail -3 gpurun_out/ab.err
