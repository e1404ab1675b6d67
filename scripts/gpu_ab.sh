#!/bin/bash
set -u
mkdir -p gpurun_out
for g in 1; do for cps in 3 4 5 6; do for ch in 2 3 4; do
  PCGRL_HOST_GRAPH=$g PCGRL_INC_CPS=$cps PCGRL_HOST_CHUNKS=$ch timeout 100 python scripts/exp_chunked_device.py 2>>gpurun_out/ab.err
done; done; done | tee gpurun_out/r02l_chunked_device.txt
PCGRL_HOST_GRAPH=0 PCGRL_INC_CPS=4 PCGRL_HOST_CHUNKS=3 timeout 100 python scripts/exp_chunked_device.py 2>>gpurun_out/ab.err | tee -a gpurun_out/r02l_chunked_device.txt
tail -3 gpurun_out/ab.err
