#!/bin/bash
# one ncu --set full capture of rk_quad_kernel (RK4) on the C4 slab
P=${1:-rkq}
mkdir -p gpurun_out
O=gpurun_out/$P
timeout 150 ncu --set full --clock-control none --import-source on -k regex:rk_quad_kernel -s 1 -c 1 -f -o ${O}_ncu_rk_quad_slab \
    python tools/quick_bench.py --spin 0 --reps 1 --steps 3 --method 1 > ${O}_ncu_rk_quad_slab.log 2>&1
tail -2 ${O}_ncu_rk_quad_slab.log
