#!/bin/bash
# One GPU-box round: parity tests, bench lines of every workload, reference arm, ncu launch list of the headline.
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
timeout 300 python bench.py --steps 800 --warmup 10 > gpurun_out/bench_binary-narrow-16x16.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_binary-narrow-16x16.json
for wl in binary-wide-ctrl-16x16 binary-turtle-16x16 zelda-turtle-7x11 zelda-narrow-7x11; do
  timeout 200 python bench.py --steps 400 --warmup 10 --workload $wl --no-cpu-baseline > gpurun_out/bench_$wl.json 2>> gpurun_out/bench.err
done
for wl in minecraft_3D_maze-narrow-14x14x14 sokoban-cellular-5x5 sokoban-narrow-5x5 smb-narrow-116x16; do
  timeout 300 python bench.py --workload $wl --steps 40 --warmup 4 --cpu-seconds 6 > gpurun_out/bench_$wl.json 2>> gpurun_out/bench.err
done
timeout 200 python bench.py --envs 65536 --steps 800 --warmup 10 --no-cpu-baseline > gpurun_out/bench_binary-narrow-16x16_65k.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2>> gpurun_out/bench.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_")[1][:-5], "value %.4g" % d["value"], "e2e %.4g" % d.get("e2e", {}).get("value", float("nan")),
              "frac %.3f" % d.get("roofline", {}).get("frac", float("nan")), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/bench.err
