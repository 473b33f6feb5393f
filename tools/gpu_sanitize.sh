#!/bin/bash
# compute-sanitizer on the tests of the fused kernels and of the Run pipeline
# (gpurun --timeout 900 -- "bash tools/gpu_sanitize.sh r02").
P=${1:-san}
mkdir -p gpurun_out; O=gpurun_out/${P}_sanitizer
S=/usr/local/cuda/bin/compute-sanitizer
# pairs: planned sub-cycling, lazy clip redo, nonzero minima, K=40 masked tile; chains; the chunk-major Run, static
# import fields, rejected attempts inside a Run
T1='tests/test_gpu_fusion.py -k "lazy or (bit_identical and 10) or variants or dt_min or k40 or (subcycling_regime and pairs) or (layer_counts and (33 or 40 or 63 or 64))"'
T2='tests/test_gpu_coupling.py tests/test_gpu_component.py -k "static_import or zero_copy or (chunk_major and 4-3600) or rejected_attempt or (run_exchange_equals and 4-3600) or export or cadence"'
timeout 300 bash -c "$S --tool memcheck --error-exitcode 3 python -m pytest $T1 -m gpu -q -x" > ${O}_memcheck_fusion.log 2>&1
echo "memcheck fusion rc=$?" >> ${O}_memcheck_fusion.log
timeout 200 bash -c "$S --tool memcheck --error-exitcode 3 python -m pytest $T2 -m gpu -q -x" > ${O}_memcheck_exchange.log 2>&1
echo "memcheck exchange rc=$?" >> ${O}_memcheck_exchange.log
timeout 200 bash -c "$S --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_fusion.py -k 'lazy or (bit_identical and pairs and 10-2)' -m gpu -q -x" > ${O}_racecheck_pairs.log 2>&1
echo "racecheck rc=$?" >> ${O}_racecheck_pairs.log
timeout 200 bash -c "$S --tool initcheck --error-exitcode 3 python -m pytest $T2 -m gpu -q -x" > ${O}_initcheck_exchange.log 2>&1
echo "initcheck rc=$?" >> ${O}_initcheck_exchange.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" ${O}_*.log
