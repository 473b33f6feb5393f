#!/bin/bash
# GPU session for rk_quad_kernel: the Runge-Kutta fusion tests, then RK4 / RK4-3/8 on the C4 slab with 4 and 2 stages per launch
P=${1:-rkq}
mkdir -p gpurun_out
O=gpurun_out/$P
timeout 400 python -m pytest tests/test_gpu_fusion.py -m gpu -q -x -k "rk_" > ${O}_pytest_rk.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest_rk.log
tail -5 ${O}_pytest_rk.log
for st in 4 2; do for m in 1 3; do
  timeout 90 python tools/quick_bench.py --spin 0 --reps 2 --steps 4 --method $m --rk-stages $st 2>&1 | grep rep | sed "s/^/stages $st method $m /"
done; done | tee ${O}_rk_timings.log
