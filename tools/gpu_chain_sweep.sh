#!/bin/bash
# chain_kernel variants on one GPU box: fusion tests on the shipped build, then timings of every build
# under build/ (compile-time variants) on the C2 tile, a 316x316 tile and the C3 tile.
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_coupling.py "tests/test_gpu_parity.py::test_one_step_parity" \
    -m gpu -q -x > $O/pytest_fused2.log 2>&1
echo "pytest rc=$?" >> $O/pytest_fused2.log
tail -3 $O/pytest_fused2.log
QB="python tools/quick_bench.py --spin 0 --reps 3 --steps 32 --fusion chains"
{
  for lib in mossco_code_b200/libmsed_b200.so build/*.so; do
    echo "== $lib C2 100x100x30";  MSED_LIB=$lib timeout 60 $QB --inum 100 --jnum 100 --knum 30 | grep rep2
    echo "== $lib 316x316x30";     MSED_LIB=$lib timeout 60 $QB --inum 316 --jnum 316 --knum 30 | grep rep2
    echo "== $lib C3 1000x1000x30 45% land"; MSED_LIB=$lib timeout 90 $QB --inum 1000 --jnum 1000 --knum 30 --land 0.45 | grep rep2
  done
  echo "== pairs C3"; timeout 90 python tools/quick_bench.py --spin 0 --reps 3 --steps 32 --fusion pairs --inum 1000 --jnum 1000 --knum 30 --land 0.45 | grep rep2
  for n in 150 200 250; do
    for f in chains pairs; do echo "== ${n}x${n}x30 $f"; timeout 60 python tools/quick_bench.py --spin 0 --reps 3 --steps 32 --fusion $f --inum $n --jnum $n --knum 30 | grep rep2; done
  done
} > $O/qb_chain2.log 2>&1
cat $O/qb_chain2.log
timeout 120 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 1 -c 1 -f -o $O/ncu_chain_c3_v2 \
    $QB --inum 1000 --jnum 1000 --knum 30 --land 0.45 --reps 1 > $O/ncu_chain_c3_v2.log 2>&1
