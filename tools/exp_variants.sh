#!/bin/bash
# kbench (session mode: classify + flux per call, 4320x2160) of the default build and of experiment variants
#   tools/exp_variants.sh <tag> <variant> [...]
TAG=$1; shift
out=gpurun_out/exp_${TAG}.txt
: > $out
for lib in default "$@"; do
  if [ "$lib" = default ]; then unset AEROBULK_GPU_LIB; else export AEROBULK_GPU_LIB=$PWD/aerobulk_b200/build/libaerobulk_gpu_$lib.so; fi
  echo "=== lib=$lib" >> $out
  KBENCH_QUICK=${KBENCH_QUICK-1} python tools/kbench.py 4320 2160 2>&1 | grep -v "^lib" >> $out
done
cat $out
