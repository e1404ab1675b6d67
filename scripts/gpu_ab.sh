#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02m_bench_steps20.json 2> gpurun_out/r02m.err; tail -2 gpurun_out/r02m.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02m_bench_steps20.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g frac %.3f kernel_ms %.4f launches %d" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["gpu_launches"]))
print(d["roofline"]["kernel"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["value"])
PY
