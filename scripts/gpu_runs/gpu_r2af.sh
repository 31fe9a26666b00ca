#!/bin/bash
# round-2 GPU call AF: short machine slices as streaming ticks at low occupancy (threshold x pops per warp)
OUT=gpurun_out
mkdir -p $OUT
for cfg in "20000 512" "5000 256" "40000 1024" "20000 128"; do
  set -- $cfg
  GLC_STREAM_MACHINE_ABOVE=$1 GLC_STREAM_MACHINE_BUDGET=$2 GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2af_forest.log 2>&1; echo "machine above $1, $2 pops per warp: exit $?"
  grep "FOREST\|forest async" $OUT/r2af_forest.log | tail -2 | cut -c1-200
done
