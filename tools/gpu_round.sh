#!/bin/bash
# One GPU-box session, most important first (the box budget may end it early; every stage writes its own
# files under gpurun_out/): tests of the fused kernels, chain/pair/single kernel timings on the configs
# with knum <= 32, bench lines, ncu captures of chain_kernel, then the rest of the GPU test suite.
# usage: gpurun --timeout 900 -- 'bash tools/gpu_round.sh'
mkdir -p gpurun_out
O=gpurun_out
date +%s > $O/t_start.txt
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1

# ---- 1. tests that exercise the fused kernels ---------------------------------------------------------
timeout 300 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_coupling.py \
    "tests/test_gpu_parity.py::test_one_step_parity" tests/test_gpu_edge.py -m gpu -q --durations=8 \
    > $O/pytest_fused.log 2>&1
echo "pytest rc=$?" >> $O/pytest_fused.log
tail -4 $O/pytest_fused.log

# ---- 2. kernel timings: chains vs pairs vs single steps ---------------------------------------------
QB="python tools/quick_bench.py --spin 0 --reps 3 --steps 32"
{
  for f in chains pairs off; do echo "== C2 100x100x30 fusion=$f"; timeout 60 $QB --inum 100 --jnum 100 --knum 30 --fusion $f; done
  echo "== C2 100x100x30 fusion=chains, 3 CTAs/SM build (80 registers)"
  MSED_LIB=build/libmsed_mb3.so timeout 60 $QB --inum 100 --jnum 100 --knum 30 --fusion chains
  for f in chains pairs; do echo "== C3 1000x1000x30 45% land fusion=$f"; timeout 90 $QB --inum 1000 --jnum 1000 --knum 30 --land 0.45 --fusion $f; done
  echo "== C3 fusion=chains, 3 CTAs/SM build"
  MSED_LIB=build/libmsed_mb3.so timeout 90 $QB --inum 1000 --jnum 1000 --knum 30 --land 0.45 --fusion chains
  for f in chains pairs; do echo "== 316x316x30 fusion=$f"; timeout 60 $QB --inum 316 --jnum 316 --knum 30 --fusion $f; done
} > $O/qb_chain.log 2>&1
grep -E "^==|rep2" $O/qb_chain.log

# ---- 3. bench lines -------------------------------------------------------------------------------
timeout 120 python bench.py --workload c2 --steps 200 --warmup 5 > $O/bench_c2.json 2> $O/bench_c2.err
timeout 150 python bench.py --workload c3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
cat $O/bench_c2.json $O/bench_c3.json | cut -c1-300

# ---- 4. ncu: full capture of one chain launch (C3 tile, C2 tile), launch list of the C2 bench ---------
timeout 150 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 1 -c 1 -f -o $O/ncu_chain_c3 \
    $QB --inum 1000 --jnum 1000 --knum 30 --land 0.45 --fusion chains --reps 1 > $O/ncu_chain_c3.log 2>&1
timeout 90 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 1 -c 1 -f -o $O/ncu_chain_c2 \
    $QB --inum 100 --jnum 100 --knum 30 --fusion chains --reps 1 > $O/ncu_chain_c2.log 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_c2.csv \
    python bench.py --workload c2 --steps 40 --warmup 3 --no-cpu-baseline > $O/ncu_c2.log 2>&1
date +%s > $O/t_stage4.txt

# ---- 5. the rest of the GPU suite, then the remaining bench lines -------------------------------------
timeout 420 python -m pytest tests -m gpu -x -q --durations=15 \
    --deselect tests/test_gpu_fusion.py --deselect tests/test_gpu_coupling.py --deselect tests/test_gpu_edge.py \
    > $O/pytest_rest.log 2>&1
echo "pytest rc=$?" >> $O/pytest_rest.log
tail -4 $O/pytest_rest.log
timeout 200 python bench.py --workload c5 --no-cpu-baseline > $O/bench_c5.json 2> $O/bench_c5.err
timeout 200 python bench.py --workload c5 --no-cpu-baseline --fusion pairs > $O/bench_c5_pairs.json 2>> $O/bench_c5.err
timeout 150 python bench.py --workload c3 --no-cpu-baseline --fusion pairs > $O/bench_c3_pairs.json 2>> $O/bench_c3.err
timeout 120 python bench.py --workload c2 --steps 200 --warmup 5 --fusion pairs --no-cpu-baseline > $O/bench_c2_pairs.json 2>> $O/bench_c2.err
date +%s > $O/t_end.txt
ls -la $O
