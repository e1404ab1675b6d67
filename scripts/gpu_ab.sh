#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest_gpu.txt
(PCGRL_HOST_TRACE=2 timeout 200 python scripts/host_trace.py 2>&1 | grep "host trace\|call 4[0-9]"; python scripts/pcie_probe.py 2>&1 | tail -6) | tee gpurun_out/r02_host_trace.txt
timeout 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_final.err
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_steps20.json 2>> gpurun_out/r02_final.err
python - <<'PY'
import json
for f in ("r02_bench", "r02_bench_steps20"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, "value %.4g e2e %.4g frac %.3f kernel_ms %.4f launches %d clocks %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["gpu_launches"], d["clocks"]))
    for k, v in d["configs"].items():
        print("   ", k, v.get("error") or "value %.4g e2e %.4g" % (v["value"], v["e2e"]["value"]))
PY
tail -3 gpurun_out/r02_final.err
