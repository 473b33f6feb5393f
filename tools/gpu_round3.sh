#!/bin/bash
# GPU session of the build with rk_quad_kernel: tools/gpu_round.sh, then compute-sanitizer (memcheck, racecheck) on the
# Runge-Kutta fusion tests in mode quad, then RK timings with 4 and 2 stages per launch.
P=${1:-r02_final3}
bash tools/gpu_round.sh $P
O=gpurun_out/${P}
S=/usr/local/cuda/bin/compute-sanitizer
T='tests/test_gpu_fusion.py -k "rk_ and quad and (bit_identical and (5 or 7 or 20) or variants or nan)"'
timeout 240 bash -c "$S --tool memcheck --error-exitcode 3 python -m pytest $T -m gpu -q -x" > ${O}_sanitizer_memcheck_rk_quad.log 2>&1
echo "memcheck rk quad rc=$?" >> ${O}_sanitizer_memcheck_rk_quad.log
timeout 200 bash -c "$S --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_fusion.py -k 'rk_ and quad and bit_identical and 7' -m gpu -q -x" > ${O}_sanitizer_racecheck_rk_quad.log 2>&1
echo "racecheck rk quad rc=$?" >> ${O}_sanitizer_racecheck_rk_quad.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" ${O}_sanitizer_*rk_quad.log
for st in 4 2; do for m in 1 3; do
  timeout 90 python tools/quick_bench.py --spin 0 --reps 2 --steps 4 --method $m --rk-stages $st 2>&1 | grep rep | sed "s/^/stages $st method $m /"
done; done > ${O}_rk_timings.log
cat ${O}_rk_timings.log
