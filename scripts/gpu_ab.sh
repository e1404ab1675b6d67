#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "smb or trace or fixtures or search" 2>&1 | tail -5
for n in 65536 262144 1048576; do
timeout 300 python bench.py --workload smb-narrow-116x16 --envs $n --steps 20 --warmup 3 --no-cpu-baseline --no-configs 2>>gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('smb $n envs: value %.4g e2e %.4g kernel_ms %.3f launches %d' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_launch'], d['gpu_launches']))"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 30 --csv --log-file gpurun_out/smb_launches.csv python bench.py --workload smb-narrow-116x16 --steps 12 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > /dev/null 2>>gpurun_out/ab.err
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/smb_launches.csv")) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:60]].append(float(r[vi].replace(",", "")))
    except ValueError: pass
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-62s n=%3d avg_us=%9.1f" % (k, len(v), sum(v) / len(v) / 1e3))
PY
tail -3 gpurun_out/ab.err
