#!/bin/bash
# end-of-round-2 evidence on the final tree: parity log, bench lines (default and the driver's 20/5), reference arm, ncu
# launch list of the headline command, memcheck of the kernels added last.  (The ncu --set full summaries and the
# observation-writer lines under profiles/ are of kernels this tree did not change.)
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/r02_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_final.err
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_steps20.json 2>> gpurun_out/r02_final.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_final.err
python - <<'PY'
import json
for f in ("r02_bench", "r02_bench_steps20"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, "value %.4g e2e %.4g frac %.3f kernel_ms %.4f launches %d clocks %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["gpu_launches"], d["clocks"]))
    for k, v in d["configs"].items():
        print("   ", k, v.get("error") or "value %.4g e2e %.4g frac %.4f kernel_ms %.3f %s" % (v["value"], v["e2e"]["value"], v["roofline"]["frac"], v["roofline"]["kernel_ms_per_launch"], v["roofline"]["kernel"][:40]))
d = json.loads(open("gpurun_out/r02_bench_reference.json").read().strip().splitlines()[-1])
print("reference arm", d["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"], {k: round(v["value"]) for k, v in d["configs"].items()})
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/r02_launches_binary_narrow.csv \
    python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r02_ncu_launch.log 2>&1
python - <<'PY' | tee gpurun_out/r02_launch_shares.txt
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_launches_binary_narrow.csv")) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:70]].append(float(r[vi].replace(",", "")))
    except ValueError: pass
tot = sum(sum(v) for v in agg.values())
print("ncu launch list of `bench.py --steps 60 --warmup 5 --no-e2e --no-cpu-baseline --no-configs` (launches 40..160)")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-72s n=%3d avg_us=%8.1f share=%5.1f%%" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
PY
MEMCHECK_ONLY=new timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/memcheck_small.py > gpurun_out/r02_memcheck_new_kernels.txt 2>&1; echo "memcheck rc=$?"; tail -12 gpurun_out/r02_memcheck_new_kernels.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
tail -5 gpurun_out/r02_final.err
