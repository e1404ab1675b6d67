#!/bin/bash
# round-2 GPU call D: full parity suite (new tests), sokoban after the job split, bench
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r02d_pytest_gpu.txt
for wl in sokoban-cellular-5x5 sokoban-narrow-5x5; do
  for envs in 1048576 4194304; do
    echo -n "$wl envs=$envs: "
    timeout 300 python bench.py --workload $wl --envs $envs --steps 20 --warmup 3 --no-cpu-baseline 2>>gpurun_out/r02d.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.4g e2e %.4g kernel_ms %.3f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_launch']))"
  done
done | tee gpurun_out/r02d_sokoban.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 40 --csv --log-file gpurun_out/r02d_launches_sokoban.csv \
    python bench.py --workload sokoban-cellular-5x5 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02d_ncu_sok.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02d_launches_sokoban.csv")) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:70]].append(float(r[vi].replace(",", "")))
    except ValueError: pass
for k, v in agg.items(): print(k, "n=%d avg_us=%.1f max_us=%.1f" % (len(v), sum(v) / len(v) / 1e3, max(v) / 1e3))
PY
timeout 600 python bench.py > gpurun_out/r02d_bench.json 2>> gpurun_out/r02d.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02d_bench.json").read().strip().splitlines()[-1])
print("headline value %.4g e2e %.4g frac %.3f kernel_ms %.4f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"]))
for k, v in d["configs"].items():
    if "error" in v: print(k, v); continue
    print(k, "value %.4g e2e %.4g frac %.4f kernel_ms %.3f cpu %.4g" % (v["value"], v["e2e"]["value"], v["roofline"]["frac"], v["roofline"]["kernel_ms_per_launch"], v.get("cpu_baseline", {}).get("value", float("nan"))))
PY
tail -5 gpurun_out/r02d.err
