#!/bin/bash
# A/B of library variants on one box: quick_bench on the slab (fresh regime) for every build/variants/*.so and the
# in-tree library, interleaved so that all see the same clocks.  usage: gpurun -- 'bash tools/gpu_ab.sh [quick_bench args]'
mkdir -p gpurun_out; O=gpurun_out
for rep in 1 2; do
  for lib in build/variants/*.so mossco_code_b200/libmsed_b200.so; do
    echo "== $lib (rep $rep)"
    MSED_LIB=$PWD/$lib timeout 120 python tools/quick_bench.py --spin 4 --reps 3 --steps 20 "$@" 2>&1 | grep rep
  done
done | tee $O/ab.log
