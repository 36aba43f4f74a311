#!/bin/bash
# Tuning sweep of the fused assembly kernel: CTA size (compile-time) x tile budget (FES_TILE_INC).
# Rebuilds assemble.cu on the GPU box for each CTA size; restores the default build at the end.
cd "$(dirname "$0")/.."
for T in ${THREADS:-128 256}; do
  touch finite-element-solver_b200/csrc/assemble.cu
  FES_NVCC_EXTRA="-DFES_TILE_THREADS=$T" python finite-element-solver_b200/build.py > /dev/null
  for INC in ${INCS:-192 256 320 384 512}; do
    echo "== threads=$T inc=$INC"
    FES_TILE_INC=$INC python tools/time_assembly.py ${CASES:-c2,tet,c3} ${TETN:-101} 2>&1 | grep -v "vector mass\|Laplacian" | cut -c1-${CUT:-120}
  done
done
touch finite-element-solver_b200/csrc/assemble.cu
python finite-element-solver_b200/build.py > /dev/null
