#!/bin/bash
# One-box tuning sweep: tile budget x CTA size x coordinate-slot colouring, per case.
#   SPECS="case:inc1,inc2:threads1,threads2 ..."  (threads 0 = the form's default)
cd "$(dirname "$0")/.."
OUT=${OUT:-gpurun_out/sweep.txt}
: > $OUT
run() {  # case inc threads extra-env
  local line
  line=$(env FES_PLAN_DEBUG=1 FES_TILE_INC=$2 ${3:+FES_TILE_THREADS=$3} $4 python tools/time_assembly.py $1 ${TETN:-101} 2>&1 | \
    grep -v "retry" | grep "elastic\|N=\|smem" | grep -v "c=1 \|Laplacian" | \
    sed -e 's/.*tiles=\([0-9]*\).*inc_max=\([0-9]*\) rec_max=\([0-9]*\).*nod_max=\([0-9]*\).*/tiles=\1 inc=\2 rec=\3 nod=\4/' \
        -e 's/.*"ms": \([0-9.]*\).*"hbm_frac": \([0-9.]*\).*/ms=\1 frac=\2/' -e 's/\[fes assemble\] //' | tr '\n' ' ')
  echo "$1 inc=$2 threads=${3:-default} $4: $line" | tee -a $OUT
}
for SPEC in ${SPECS:-c2:288,304,320,336:128 tet:192,256,320,384,512:128,256 c3:64,96,128:128,256}; do
  IFS=: read CASE INCS THREADS <<< "$SPEC"
  for INC in ${INCS//,/ }; do
    for T in ${THREADS//,/ }; do
      run $CASE $INC $T ""
    done
  done
done
run c2 336 128 FES_PLAN_NO_COLOUR=1
run tet 512 128 FES_PLAN_NO_COLOUR=1
