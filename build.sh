#!/usr/bin/env bash
# Build the C-ABI CUDA library in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
SRC=motioncraft_b200/csrc
OUT=motioncraft_b200/libmcm_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v"
mkdir -p build
for f in gemm_tc fused_block elementwise context timing handoff pathb; do
  if [ ! -f build/$f.o ] || [ $SRC/$f.cu -nt build/$f.o ] || [ -n "$(find $SRC include -newer build/$f.o -name '*.cuh' -o -newer build/$f.o -name '*.h')" ]; then
    $NVCC $FLAGS -c $SRC/$f.cu -o build/$f.o 2> build/$f.log || { cat build/$f.log; exit 1; }
  fi
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT build/gemm_tc.o build/fused_block.o build/elementwise.o build/context.o build/timing.o build/handoff.o build/pathb.o -lcudart
echo "built $OUT"
