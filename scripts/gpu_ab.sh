#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_split.py tests/test_gpu_parity.py -x -q -k "split or pipelined or compact or paths or big" 2>&1 | tail -3
for path in inc split; do
  echo -n "path=$path: "
  PCGRL_STEP_PATH=$path timeout 200 python bench.py --steps 800 --warmup 10 --no-cpu-baseline --no-configs 2>>gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.4g e2e %.4g kernel_ms %.4f frac %.3f launches %d' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['gpu_launches']))"
done
PCGRL_STEP_PATH=inc timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/r02g_launches_inc.csv \
      python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r02g_ncu_inc.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02g_launches_inc.csv")) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:60]].append(float(r[vi].replace(",", "")))
    except ValueError: pass
for k, v in agg.items(): print(k, "n=%d avg_us=%.1f" % (len(v), sum(v) / len(v) / 1e3))
PY
tail -3 gpurun_out/ab.err
