#!/bin/bash
# quick iteration: parity tests + headline bench + ncu full capture of the step kernel
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 800 --warmup 10 --no-cpu-baseline > gpurun_out/bench_narrow_1m.json 2> gpurun_out/bench.err; cut -c1-700 gpurun_out/bench_narrow_1m.json
python bench.py --steps 400 --warmup 10 --workload zelda-turtle-7x11 --no-cpu-baseline --no-e2e > gpurun_out/bench_zelda.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_zelda.json
ncu --set full --clock-control none --import-source on -k regex:k_step_bitboard -s 8 -c 1 -f -o gpurun_out/prof_step \
    python bench.py --steps 12 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/bench.err
