#!/bin/bash
# compute-sanitizer on the tests of the fused kernels (gpurun --timeout 400 -- "bash tools/gpu_sanitize.sh").
mkdir -p gpurun_out; O=gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 170 $S --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_fusion.py -m gpu -q -x \
    -k "chain or (bit_identical and chains and 10) or variants or dt_min" > $O/sanitizer_memcheck_chain.log 2>&1
echo "memcheck rc=$?" >> $O/sanitizer_memcheck_chain.log
timeout 120 $S --tool initcheck --error-exitcode 3 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_coupling.py -m gpu -q -x \
    -k "(bit_identical and chains and 11) or soil_pelagic or (run_exchange_equals and 0-)" > $O/sanitizer_initcheck_chain.log 2>&1
echo "initcheck rc=$?" >> $O/sanitizer_initcheck_chain.log
grep -E "ERROR SUMMARY|passed|failed|rc=" $O/sanitizer_memcheck_chain.log $O/sanitizer_initcheck_chain.log
