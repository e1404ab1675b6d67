#!/bin/bash
# N-GPU bench line (one rank per GPU over NCCL), as the driver launches it
set -u
N=$1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 400 --warmup 10 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
tail -3 gpurun_out/r02_bench_${N}gpu.err
python - $N <<'PY'
import json, sys
N = sys.argv[1]
d = json.loads(open(f"gpurun_out/r02_bench_{N}gpu.json").read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "headline value %.4g e2e %.4g frac %.3f kernel_ms %.4f bound_cores %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["config"]["host_cores_bound_rank0"]))
for k, v in d["configs"].items():
    if "error" in v: print(k, v); continue
    print(k, "value %.4g e2e %.4g kernel_ms %.3f" % (v["value"], v["e2e"]["value"], v["roofline"]["kernel_ms_per_launch"]))
PY
