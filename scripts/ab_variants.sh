#!/bin/bash
# Build kernel variants (-D knobs) here, time each on the GPU box.
#   local:  bash scripts/ab_variants.sh build "A:-DX=1" "B:-DX=0"     -> gpurun_out/variants/libA.so ...
#   remote: bash scripts/ab_variants.sh run [bench args]
set -u
cd "$(dirname "$0")/.."
if [ "$1" = build ]; then
  shift; rm -rf gpurun_variants; mkdir -p gpurun_variants
  for spec in "$@"; do
    name="${spec%%:*}"; flags="${spec#*:}"
    ( cd control_pcgrl_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared $flags \
        -o ../../gpurun_variants/lib$name.so api.cu step_bitboard.cu observe.cu ) && echo "built $name ($flags)"
  done
else
  shift
  mkdir -p gpurun_out
  for lib in gpurun_variants/lib*.so; do
    name=$(basename $lib .so)
    for rep in 1 2; do
      PCGRL_B200_LIB=$PWD/$lib timeout 90 python bench.py --no-cpu-baseline --no-e2e "$@" 2>>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', 'value %.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'kernel_ms %.4f'%d['roofline']['kernel_ms_per_launch'])"
    done
  done | tee gpurun_out/ab.txt
fi
