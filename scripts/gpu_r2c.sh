#!/bin/bash
# round-2 GPU call C: parity of the fixed split path, per-kernel times of the three step paths, ncu of the searches
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r02c_pytest_gpu.txt
for path in fused split inc; do
  PCGRL_STEP_PATH=$path timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/r02c_launches_$path.csv \
      python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r02c_ncu_$path.log 2>&1
  python - "$path" <<'PY'
import csv, sys, collections
path = sys.argv[1]
rows = [r for r in csv.reader(open(f"gpurun_out/r02c_launches_{path}.csv")) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:60]].append(float(r[vi].replace(",", "")))
    except ValueError: pass
for k, v in agg.items(): print(path, k, "n=%d avg_us=%.1f" % (len(v), sum(v) / len(v) / 1e3))
PY
done | tee gpurun_out/r02c_kernel_times.txt
for path in split inc; do
  PCGRL_STEP_PATH=$path timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_split -s 60 -c 3 -o gpurun_out/r02c_prof_$path -f \
      python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r02c_ncufull_$path.log 2>&1
  python scripts/ncu_summary.py gpurun_out/r02c_prof_$path.ncu-rep > gpurun_out/r02c_summary_$path.txt 2>&1
done
grep -A30 "k_split_stats" gpurun_out/r02c_summary_split.txt | head -40
grep -A30 "k_split_stats" gpurun_out/r02c_summary_inc.txt | head -40
