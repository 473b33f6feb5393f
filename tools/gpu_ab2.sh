#!/bin/bash
# A/B on one box: every library in build/variants plus the in-tree one, each with both feeds of pair_kernel
# (MSED_PAIR_FEED=cpasync | default), quick_bench on the slab.  usage: gpurun -- 'bash tools/gpu_ab2.sh'
mkdir -p gpurun_out
for rep in 1 2; do
  for lib in build/variants/*.so mossco_code_b200/libmsed_b200.so; do
    for feed in cpasync bulk; do
      echo "== $lib feed=$feed (rep $rep)"
      MSED_LIB=$PWD/$lib MSED_PAIR_FEED=$feed timeout 120 python tools/quick_bench.py --spin 4 --reps 2 --steps 20 "$@" 2>&1 | grep rep
    done
  done
done | tee gpurun_out/ab2.log
