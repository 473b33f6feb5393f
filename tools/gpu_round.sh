#!/bin/bash
# One GPU-box session for the build of a round: the GPU test suite, smoke(), the bench lines of every workload,
# the ncu launch list and full captures of the fused kernels.  Everything lands in gpurun_out/<prefix>_*.
# usage: gpurun --timeout 1200 -- 'bash tools/gpu_round.sh r02_final'
P=${1:-round}
mkdir -p gpurun_out
O=gpurun_out/$P
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > ${O}_gpu.txt 2>&1

timeout 500 python -m pytest tests -m gpu -q --durations=8 > ${O}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> ${O}_pytest_gpu.log
tail -3 ${O}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; tail -1 ${O}_smoke.log

timeout 400 python bench.py > ${O}_bench_c4.json 2> ${O}_bench_c4.err
timeout 200 python bench.py --impl reference --steps 20 --warmup 3 > ${O}_bench_c4_reference_arm.json 2>> ${O}_bench_c4.err
timeout 150 python bench.py --workload c2 --steps 200 --warmup 5 --no-cpu-baseline --tts-days 10 > ${O}_bench_c2.json 2> ${O}_bench_c2.err
timeout 150 python bench.py --workload c3 --no-cpu-baseline > ${O}_bench_c3.json 2> ${O}_bench_c3.err
timeout 200 python bench.py --workload c5 --no-cpu-baseline > ${O}_bench_c5.json 2> ${O}_bench_c5.err
timeout 150 python bench.py --workload c4slab --no-cpu-baseline > ${O}_bench_c4slab.json 2> ${O}_bench_c4slab.err
timeout 150 python bench.py --workload c4slab --regime subcycling --steps 50 --no-cpu-baseline > ${O}_bench_c4slab_subcycling.json 2>> ${O}_bench_c4slab.err
timeout 150 python bench.py --workload c4slab --fusion off --no-cpu-baseline > ${O}_bench_c4slab_nofusion.json 2>> ${O}_bench_c4slab.err
for f in ${O}_bench_c4.json ${O}_bench_c2.json ${O}_bench_c3.json ${O}_bench_c5.json ${O}_bench_c4slab.json ${O}_bench_c4slab_subcycling.json; do cut -c1-170 $f; done

QB="python tools/quick_bench.py --spin 0 --reps 1 --steps 32"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches_c4slab.csv \
    python bench.py --workload c4slab --steps 10 --warmup 3 --no-cpu-baseline > ${O}_ncu_launches.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 2 -c 1 -f -o ${O}_ncu_pair_slab \
    $QB > ${O}_ncu_pair_slab.log 2>&1
timeout 90 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 1 -c 1 -f -o ${O}_ncu_chain_c2 \
    $QB --inum 100 --jnum 100 --knum 30 > ${O}_ncu_chain_c2.log 2>&1
for m in 1 3; do timeout 60 python tools/quick_bench.py --spin 0 --reps 2 --steps 4 --method $m 2>&1 | grep rep | sed "s/^/method $m /"; done > ${O}_rk_timings.log
ls -la gpurun_out | grep $P
