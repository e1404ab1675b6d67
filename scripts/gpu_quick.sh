#!/bin/bash
# quick GPU check: the whole parity suite
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_quick.txt
