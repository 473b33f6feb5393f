#!/bin/bash
# last check of a build: the GPU suite, smoke, the C3-shaped quick timing and the C3 / C4 bench lines
P=${1:-final}
mkdir -p gpurun_out
O=gpurun_out/$P
timeout 500 python -m pytest tests -m gpu -q > ${O}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest_gpu.log
tail -3 ${O}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; tail -1 ${O}_smoke.log
timeout 90 python tools/quick_bench.py --inum 1000 --jnum 1000 --knum 30 --land 0.45 --spin 0 --reps 3 --steps 20 2>&1 | grep rep | tee ${O}_c3_quick.log
timeout 150 python bench.py --workload c3 --no-cpu-baseline > ${O}_bench_c3.json 2> ${O}_bench_c3.err
timeout 400 python bench.py > ${O}_bench_c4.json 2> ${O}_bench_c4.err
for f in ${O}_bench_c3.json ${O}_bench_c4.json; do cut -c1-200 $f; done
