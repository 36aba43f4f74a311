#!/bin/bash
# Tuning sweep of the tile record budget (FES_TILE_INC) per case: CASE:inc1,inc2,...
cd "$(dirname "$0")/.."
for SPEC in ${SPECS:-c2:288,320,336,352,384 c3:96,128,160,192,224 tet:448,512,576}; do
  CASE=${SPEC%%:*}; INCS=${SPEC#*:}
  for INC in ${INCS//,/ }; do
    echo -n "$CASE inc=$INC: "
    FES_PLAN_DEBUG=1 FES_TILE_INC=$INC python tools/time_assembly.py $CASE ${TETN:-101} 2>&1 | grep -v "mass\|Laplacian\|retry" | \
      sed -e 's/.*tiles=\([0-9]*\).*inc_max=\([0-9]*\) rec_max=\([0-9]*\).*item_max=\([0-9]*\).*items=\([0-9]*\).*/tiles=\1 inc=\2 rec=\3 imax=\4 items=\5/' -e 's/.*"ms": \([0-9.]*\).*/ms=\1/' | head -2 | tr '\n' ' '
    echo
  done
done
