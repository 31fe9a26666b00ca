#!/bin/bash
# round-2 GPU call AL: 64-thread blocks x 4 per SM for the warp-synchronous kernels (fewer warps meet at each block barrier)
OUT=gpurun_out
mkdir -p $OUT
for lib in libglcb200_b64.so libglcb200.so; do
  GLC_LIB_PATH=$PWD/galacticus_b200/$lib timeout 300 python scripts/knobs.py 1000000 | grep KNOBS
  GLC_LIB_PATH=$PWD/galacticus_b200/$lib GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2al_forest.log 2>&1; echo "$lib exit $?"
  grep "FOREST\|forest async" $OUT/r2al_forest.log | tail -3 | cut -c1-200
done
