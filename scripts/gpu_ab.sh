#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
PCGRL_STEP_PATH=inc timeout 200 python bench.py --steps 800 --warmup 10 --no-cpu-baseline --no-configs 2>>gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('inc: value %.4g e2e %.4g kernel_ms %.4f frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac']))"
timeout 300 python scripts/bench_observe.py > gpurun_out/r02k_observe.jsonl 2>> gpurun_out/ab.err; cut -c1-200 gpurun_out/r02k_observe.jsonl
timeout 300 python scripts/bench_rl_loop.py > gpurun_out/r02k_rl_loop.jsonl 2>> gpurun_out/ab.err; cat gpurun_out/r02k_rl_loop.jsonl
tail -3 gpurun_out/ab.err
