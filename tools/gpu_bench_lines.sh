#!/bin/bash
# The bench lines of every workload on one GPU box (gpurun --timeout 600 -- "bash tools/gpu_bench_lines.sh").
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python bench.py > $O/bench_c4.json 2> $O/bench_c4.err
timeout 120 python bench.py --workload c2 --steps 200 --warmup 5 --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
timeout 150 python bench.py --workload c3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
timeout 200 python bench.py --workload c5 --no-cpu-baseline > $O/bench_c5.json 2> $O/bench_c5.err
timeout 150 python bench.py --workload c4slab --no-cpu-baseline > $O/bench_c4slab.json 2> $O/bench_c4slab.err
timeout 120 python bench.py --workload c2 --steps 200 --warmup 5 --no-cpu-baseline --fusion pairs > $O/bench_c2_pairs.json 2>> $O/bench_c2.err
timeout 120 python bench.py --workload c2 --steps 200 --warmup 5 --no-cpu-baseline --fusion off > $O/bench_c2_off.json 2>> $O/bench_c2.err
timeout 150 python bench.py --workload c3 --no-cpu-baseline --fusion chains > $O/bench_c3_chains.json 2>> $O/bench_c3.err
timeout 300 python bench.py --steps 100 --no-cpu-baseline > $O/bench_c4_k100.json 2>> $O/bench_c4.err
for f in $O/*.err; do tail -n 2 "$f"; done
