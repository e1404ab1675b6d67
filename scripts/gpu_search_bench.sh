#!/bin/bash
# bench lines of the search-based workloads + ncu launch list / full capture of k_step_search
set -u
mkdir -p gpurun_out
for wl in minecraft_3D_maze-narrow-14x14x14 sokoban-cellular-5x5 sokoban-narrow-5x5 smb-narrow-116x16; do
  timeout 600 python bench.py --workload $wl --steps ${STEPS:-60} --warmup 5 --cpu-seconds 6 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$wl.json").read())
    print("$wl", "value %.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "cpu %.4g"%(d["cpu_baseline"]["value"]), d["mean_stats"])
except Exception as e:
    print("$wl FAILED", e); print(open("gpurun_out/bench_$wl.err").read()[-1500:])
PY
done
for wl in minecraft_3D_maze-narrow-14x14x14 smb-narrow-116x16 sokoban-cellular-5x5; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_search -s 6 -c 1 -f -o gpurun_out/prof_$wl \
     python bench.py --workload $wl --steps 10 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$wl.log 2>&1
done
ls -la gpurun_out | tail -12
