#!/bin/bash
# GPU-box run: parity tests, then the e2e (host-buffer) arm of bench.py for several chunk counts of the
# pipelined pcgrl_step_host.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for c in 1 2 4 8 16 32; do
  PCGRL_HOST_CHUNKS=$c python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/e2e_chunks_$c.json 2> gpurun_out/e2e_chunks_$c.err
  python - <<PY
import json
l = json.loads(open("gpurun_out/e2e_chunks_$c.json").read().strip().splitlines()[-1])
print("chunks", $c, "device %.3e" % l["value"], "e2e %.3e" % l["e2e"]["value"])
PY
done
python bench.py --steps 800 --warmup 10 > gpurun_out/bench_narrow_1m.json 2> gpurun_out/bench_narrow_1m.err; tail -c 2500 gpurun_out/bench_narrow_1m.json
