#!/bin/bash
# One GPU-box session for the final build of a round: the GPU test suite, smoke(), the bench lines of every
# workload, the ncu launch list and full captures of the two fused kernels.  Everything lands in gpurun_out/.
# usage: gpurun --timeout 900 -- 'bash tools/gpu_round.sh'
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/t_start.txt
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1

timeout 400 python -m pytest tests -m gpu -q --durations=10 > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log

timeout 300 python bench.py > $O/bench_c4.json 2> $O/bench_c4.err
timeout 120 python bench.py --workload c2 --steps 200 --warmup 5 --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
timeout 150 python bench.py --workload c3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
timeout 200 python bench.py --workload c5 --no-cpu-baseline > $O/bench_c5.json 2> $O/bench_c5.err
timeout 150 python bench.py --workload c4slab --no-cpu-baseline > $O/bench_c4slab.json 2> $O/bench_c4slab.err
timeout 120 python bench.py --workload c2 --steps 200 --warmup 5 --no-cpu-baseline --fusion pairs > $O/bench_c2_pairs.json 2>> $O/bench_c2.err
for f in $O/bench_c4.json $O/bench_c2.json $O/bench_c3.json $O/bench_c5.json $O/bench_c4slab.json; do cut -c1-220 $f; done
date +%s > $O/t_bench.txt

QB="python tools/quick_bench.py --spin 0 --reps 1 --steps 32"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c4slab.csv \
    python bench.py --workload c4slab --steps 10 --warmup 3 --no-cpu-baseline > $O/ncu_c4slab.log 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_c2.csv \
    python bench.py --workload c2 --steps 40 --warmup 3 --no-cpu-baseline > $O/ncu_c2.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 2 -c 1 -f -o $O/ncu_pair_slab \
    $QB > $O/ncu_pair_slab.log 2>&1
timeout 90 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 1 -c 1 -f -o $O/ncu_chain_c2 \
    $QB --inum 100 --jnum 100 --knum 30 > $O/ncu_chain_c2.log 2>&1
date +%s > $O/t_end.txt
ls -la $O
