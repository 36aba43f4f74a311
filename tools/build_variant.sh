#!/bin/bash
# Build an A-B variant of libfes_b200.so: tools/build_variant.sh NAME "-DFLAG=1 ..."  ->
# finite-element-solver_b200/build/variants/NAME.so (rides to the GPU box; select it with FES_B200_LIB=...).
set -e
cd "$(dirname "$0")/.."
NAME=$1; FLAGS=$2
P=finite-element-solver_b200
O=$P/build/variants/$NAME.objs
mkdir -p $O
for f in core plan plan_runs assemble pcg simp fields; do
  if [ "$f" = assemble ] || [ ! -f $O/$f.o ]; then
    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $FLAGS -c $P/csrc/$f.cu -o $O/$f.o &
  fi
done
wait
nvcc -shared -cudart shared -Xlinker -rpath=/usr/local/cuda/lib64 -o $P/build/variants/$NAME.so $O/*.o
echo $P/build/variants/$NAME.so
