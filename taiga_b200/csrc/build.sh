#!/bin/bash
# Builds libtaiga_b200.so in-tree for sm_100a (the .so travels to the GPU box with the repo snapshot).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v"
mkdir -p build
pids=()
for f in *.cu; do
  o=build/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find . -maxdepth 1 -name '*.cuh' -newer "$o")" ] || [ ../../include/taiga_b200.h -nt "$o" ]; then
    ( $NVCC $FLAGS -c "$f" -o "$o" > build/${f%.cu}.log 2>&1 || { cat build/${f%.cu}.log; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o ../libtaiga_b200.so build/*.o -lcudart
echo "built $(realpath ../libtaiga_b200.so)"
