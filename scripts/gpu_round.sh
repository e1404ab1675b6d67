#!/bin/bash
# One GPU-box round: parity tests, bench lines, ncu launch list + full capture of the step kernel.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 800 --warmup 10 > gpurun_out/bench_narrow_1m.json 2> gpurun_out/bench_narrow_1m.err; tail -c 3000 gpurun_out/bench_narrow_1m.json
python bench.py --steps 800 --warmup 10 --envs 65536 --no-cpu-baseline > gpurun_out/bench_narrow_65k.json 2>> gpurun_out/bench_narrow_1m.err; tail -c 1500 gpurun_out/bench_narrow_65k.json
python bench.py --steps 400 --warmup 10 --workload binary-wide-ctrl-16x16 --no-cpu-baseline > gpurun_out/bench_wide_ctrl.json 2>> gpurun_out/bench_narrow_1m.err; tail -c 1500 gpurun_out/bench_wide_ctrl.json
python bench.py --steps 400 --warmup 10 --workload zelda-turtle-7x11 --no-cpu-baseline > gpurun_out/bench_zelda.json 2>> gpurun_out/bench_narrow_1m.err; tail -c 1500 gpurun_out/bench_zelda.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step_bitboard -s 8 -c 2 -f -o gpurun_out/prof_step \
    python bench.py --steps 12 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -20
