#!/bin/bash
# round-2 GPU call A: parity tests, bench (all configs), e2e chunk sweep for the packed host path
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | head -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02a_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; tail -3 gpurun_out/r02a_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02a_bench.json").read().strip().splitlines()[-1])
print("headline value %.4g e2e %.4g frac %.3f kernel_ms %.4f cpu %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["cpu_baseline"]))
for k, v in d["configs"].items():
    if "error" in v: print(k, v); continue
    print(k, "value %.4g e2e %.4g frac %.4f kernel_ms %.3f cpu %.4g" % (v["value"], v["e2e"]["value"], v["roofline"]["frac"], v["roofline"]["kernel_ms_per_launch"], v.get("cpu_baseline", {}).get("value", float("nan"))))
PY
for ch in 1 2 4 6 8 12 16; do
  echo -n "packed chunks=$ch: "
  PCGRL_HOST_CHUNKS=$ch timeout 120 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-configs 2>>gpurun_out/r02a_bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('e2e %.4g value %.4g' % (d['e2e']['value'], d['value']))"
done | tee gpurun_out/r02a_chunk_sweep.txt
echo -n "int32 io: "; timeout 120 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-configs --int32-io 2>>gpurun_out/r02a_bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('e2e %.4g value %.4g' % (d['e2e']['value'], d['value']))" | tee -a gpurun_out/r02a_chunk_sweep.txt
timeout 300 python scripts/bench_rl_loop.py > gpurun_out/r02a_rl_loop.jsonl 2>> gpurun_out/r02a_bench.err; cat gpurun_out/r02a_rl_loop.jsonl
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 --no-configs > gpurun_out/r02a_bench_reference.json 2>> gpurun_out/r02a_bench.err; cut -c1-400 gpurun_out/r02a_bench_reference.json
tail -5 gpurun_out/r02a_bench.err
