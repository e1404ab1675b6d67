#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_split.py -x -q 2>&1 | tail -3
bash scripts/ab_variants.sh run --steps 800 --warmup 10 2>&1 | tail -60
tail -3 gpurun_out/ab.err
