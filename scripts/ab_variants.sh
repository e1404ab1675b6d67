#!/bin/bash
# Build kernel variants (-D knobs) here, time each on the GPU box.
#   local:  bash scripts/ab_variants.sh build FILE.cu "A:-DX=1" "B:-DX=0"   -> gpurun_variants/libA.so ...
#           (FILE.cu is recompiled per variant and linked with the other objects of build/obj; a name starting with
#            F_ / S_ / I_ / X_ / L_ is timed on the fused / split / split-incremental / fused-incremental / lane-group step path)
#   remote: bash scripts/ab_variants.sh run [bench args]
set -u
cd "$(dirname "$0")/.."
if [ "$1" = build ]; then
  shift; file="$1"; shift
  mkdir -p gpurun_variants
  others=$(ls build/obj/*.o | grep -v "/${file%.cu}.o")
  for spec in "$@"; do
    name="${spec%%:*}"; flags="${spec#*:}"
    ( cd control_pcgrl_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $flags \
        -c -o /tmp/ab_$name.o $file ) && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o gpurun_variants/lib$name.so /tmp/ab_$name.o $others \
        && echo "built $name ($flags)"
  done
else
  shift
  mkdir -p gpurun_out
  for lib in gpurun_variants/lib*.so; do
    name=$(basename $lib .so); name=${name#lib}
    case $name in L_*) path=lg;; F_*) path=fused;; S_*) path=split;; I_*) path=inc;; X_*) path=incfused;; *) path=incfused;; esac
    for rep in 1 2; do
      PCGRL_STEP_PATH=$path PCGRL_B200_LIB=$PWD/$lib timeout 90 python bench.py --no-cpu-baseline --no-e2e --no-configs "$@" 2>>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', '$path', 'value %.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'kernel_ms %.4f'%d['roofline']['kernel_ms_per_launch'])"
    done
  done | tee gpurun_out/ab.txt
fi
