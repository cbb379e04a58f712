#!/bin/bash
out=gpurun_out/exp_sort_${1:-r02e}.txt
: > $out
for lib in default sband075 sband1 band01 band02; do
  if [ "$lib" = default ]; then unset AEROBULK_GPU_LIB; else export AEROBULK_GPU_LIB=$PWD/aerobulk_b200/build/libaerobulk_gpu_$lib.so; fi
  for single in 1 ""; do
    echo "=== lib=$lib sort=1 single=${single:-0}" >> $out
    KBENCH_SORT=1 KBENCH_SINGLE=$single python tools/kbench.py 4320 2160 2>&1 | grep -v "^lib" >> $out
  done
done
