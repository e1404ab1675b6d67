#!/bin/bash
# GPU-box run: parity of the step paths (incl. the lane-group kernel, now the default below 12 Ki envs) and of the
# sharded vector env; lane-group kernel time at small shards; RL loop with 1..4 stream-overlapped env ranges.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_split.py tests/test_gpu_vector.py -m gpu -x -q 2>&1 | tail -8
for path in incfused lg; do
  for n in 1024 4096 16384; do
    PCGRL_STEP_PATH=$path timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_step -s 20 -c 30 --csv --log-file gpurun_out/lg_ncu_${path}_$n.csv \
      python bench.py --envs $n --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-configs > /dev/null 2>&1
    python - "$path" "$n" <<'PY'
import csv, sys
path, n = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(f"gpurun_out/lg_ncu_{path}_{n}.csv")) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
v = [float(r[vi].replace(",", "")) for r in rows[1:] if r[vi].replace(",", "").replace(".", "").isdigit()]
print("ncu", path, n, rows[1][ki][:40], "n=%d avg_us=%.1f min_us=%.1f" % (len(v), sum(v) / len(v) / 1e3, min(v) / 1e3))
PY
  done
done | tee gpurun_out/lg_ncu_times2.txt
timeout 600 python scripts/bench_rl_loop.py --obs uint8 --shards 1 2 3 4 2>&1 | tee gpurun_out/rl_loop_shards.jsonl
timeout 300 python scripts/bench_rl_loop.py --obs float32 --shards 1 2 --steps 800 2>&1 | tee -a gpurun_out/rl_loop_shards.jsonl
