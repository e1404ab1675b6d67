#!/bin/bash
# round-2 GPU call E: parity suite, observation writer (bulk store + packed u8 records), RL loop, bench
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02e_pytest_gpu.txt
timeout 300 python scripts/bench_observe.py > gpurun_out/r02e_observe.jsonl 2>> gpurun_out/r02e.err; cat gpurun_out/r02e_observe.jsonl | cut -c1-220
timeout 300 python scripts/bench_rl_loop.py > gpurun_out/r02e_rl_loop.jsonl 2>> gpurun_out/r02e.err; cat gpurun_out/r02e_rl_loop.jsonl
timeout 600 python bench.py > gpurun_out/r02e_bench.json 2>> gpurun_out/r02e.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02e_bench.json").read().strip().splitlines()[-1])
print("headline value %.4g e2e %.4g frac %.3f kernel_ms %.4f launches %d" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["gpu_launches"]))
for k, v in d["configs"].items():
    if "error" in v: print(k, v); continue
    print(k, "value %.4g e2e %.4g frac %.4f kernel_ms %.3f cpu %.4g" % (v["value"], v["e2e"]["value"], v["roofline"]["frac"], v["roofline"]["kernel_ms_per_launch"], v.get("cpu_baseline", {}).get("value", float("nan"))))
PY
tail -5 gpurun_out/r02e.err
