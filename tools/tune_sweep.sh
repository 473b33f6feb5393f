#!/bin/bash
# Builds variants of the column kernel (CTA size, CTAs/SM, ring depth) and times them with
# tools/quick_bench.py.  Development helper:  bash tools/tune_sweep.sh --build-only   (here, no GPU)
#                                             bash tools/tune_sweep.sh                (on the GPU box)
set -e
MODE="$1"
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/variants
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared"
for cfg in "128 4 4" "128 4 2" "128 4 8" "64 8 4" "256 2 4" "128 3 4" "128 5 4" "64 8 8" "32 16 8"; do
  set -- $cfg
  out=gpurun_out/variants/libmsed_b$1_m$2_s$3.so
  [ -f $out ] || $NV -DMSED_COL_BLOCK=$1 -DMSED_COL_MIN_BLOCKS=$2 -DMSED_RING_STAGES=$3 -o $out mossco_code_b200/csrc/msed.cu -ldl
done
if [ "$MODE" != "--build-only" ]; then
  for f in gpurun_out/variants/*.so; do
    echo "== $f"
    MSED_LIB=$PWD/$f python tools/quick_bench.py --steps 20 2>&1 | tail -2
  done
fi
