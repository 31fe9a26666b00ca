#!/bin/bash
# round-2 GPU call L: forest 1000 async -- tick budgets (evaluations per lane per lane pass) x spreading
OUT=gpurun_out
mkdir -p $OUT
for cfg in "1 8 4" "1 16 6" "1 4 2" "0 16 6" "0 8 4"; do
  set -- $cfg
  GLC_DRAIN_SPREAD=$1 GLC_STREAM_SPARSE_BUDGET=$2 GLC_STREAM_DENSE_BUDGET=$3 GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2l_mw1000_$1_$2_$3.log 2>&1; echo "spread=$1 sparse=$2 dense=$3 exit $?"
  grep "FOREST\|forest async" $OUT/r2l_mw1000_$1_$2_$3.log | tail -3 | cut -c1-230
done
