#!/bin/bash
# Build libnumpad_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v"
SRCS="capi.cu scan.cu assembly.cu coo.cu krylov.cu gmres.cu bicgstab.cu bandlu.cu amg.cu dist.cu spgemm.cu blocktri.cu"
mkdir -p build
OBJS=""
for f in $SRCS; do
  o=build/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ common.cuh -nt "$o" ] || [ scan.cuh -nt "$o" ] || [ spmv_tma.cuh -nt "$o" ] || [ dist.cuh -nt "$o" ]; then
    echo "nvcc $f"
    $NVCC $FLAGS -c "$f" -o "$o" 2> build/${f%.cu}.ptxas.log || { cat build/${f%.cu}.ptxas.log; exit 1; }
  fi
  OBJS="$OBJS $o"
done
$NVCC -shared -o ../libnumpad_b200.so $OBJS -lcudart -ldl
echo "built $(cd .. && pwd)/libnumpad_b200.so"
