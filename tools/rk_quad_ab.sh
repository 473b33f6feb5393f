#!/bin/bash
# rk_quad_kernel: one ncu --set full capture on the C4 slab, then RK4 timings of the library variants in build/variants
P=${1:-rkq}
mkdir -p gpurun_out
O=gpurun_out/$P
timeout 150 ncu --set full --clock-control none --import-source on -k regex:rk_quad_kernel -s 1 -c 1 -f -o ${O}_ncu_rk_quad_slab \
    python tools/quick_bench.py --spin 0 --reps 1 --steps 3 --method 1 > ${O}_ncu_rk_quad_slab.log 2>&1
for rep in 1 2; do
  for lib in build/variants/*.so mossco_code_b200/libmsed_b200.so; do
    echo "== $lib (rep $rep)"
    MSED_LIB=$PWD/$lib timeout 120 python tools/quick_bench.py --spin 0 --reps 2 --steps 4 --method 1 2>&1 | grep rep
  done
done | tee ${O}_ab.log
