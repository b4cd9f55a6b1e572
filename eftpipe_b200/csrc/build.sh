#!/bin/bash
# Build libeftb200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v"
mkdir -p build
pids=()
SRCS="api gemm layout antidiag group resum ap like producer"
for f in $SRCS; do
  $NVCC $FLAGS -c $f.cu -o build/$f.o > build/$f.log 2>&1 &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
OBJS=""
for f in $SRCS; do OBJS="$OBJS build/$f.o"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../libeftb200.so $OBJS
grep -h -E "error|spill|registers" build/*.log | grep -v " 0 bytes spill" | sort | uniq -c | sort -rn | head -20 || true
ls -la ../libeftb200.so
