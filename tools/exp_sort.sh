#!/bin/bash
# sort-policy experiment: kbench (single-call C5-like mode and 12-step session mode) for the default build and the
# band variants, auto policy (KBENCH_SORT=1) and forced (=2), at 4320x2160
out=gpurun_out/exp_sort_${1:-r02}.txt
: > $out
for lib in default band0 band01 band05; do
  for sort in 1 2; do
    for single in 1 ""; do
      if [ "$lib" = default ]; then unset AEROBULK_GPU_LIB; else export AEROBULK_GPU_LIB=$PWD/aerobulk_b200/build/libaerobulk_gpu_$lib.so; fi
      echo "=== lib=$lib sort=$sort single=${single:-0}" >> $out
      KBENCH_SORT=$sort KBENCH_SINGLE=$single python tools/kbench.py 4320 2160 2>&1 | grep -v "^lib" >> $out
    done
  done
done
cat $out
