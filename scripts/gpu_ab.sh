#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "sokoban or search or fixtures" 2>&1 | tail -3
for r in 1 2; do
timeout 300 python bench.py --workload sokoban-cellular-5x5 --steps 30 --warmup 3 --no-cpu-baseline --no-configs 2>>gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('sokoban: value %.4g e2e %.4g kernel_ms %.3f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_launch']))"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sokoban\|k_step_search -c 160 --csv --log-file gpurun_out/r02_launches_sokoban.csv python bench.py --workload sokoban-cellular-5x5 --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > /dev/null 2>>gpurun_out/ab.err
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_launches_sokoban.csv")) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:60]].append(float(r[vi].replace(",", "")) / 1e3)
    except ValueError: pass
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    v2 = sorted(v)
    print("%-62s n=%3d avg_us=%9.1f median=%9.1f p90=%9.1f max=%9.1f" % (k, len(v), sum(v) / len(v), v2[len(v2)//2], v2[int(len(v2)*0.9)], v2[-1]))
PY
tail -3 gpurun_out/ab.err
