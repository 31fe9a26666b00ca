#!/bin/bash
# round-2 GPU call AV: volume forest (configs[3], 12 500 mass-function-sampled trees): the nodes of the most massive halos one per warp
# in lane passes (priority express by halo mass) -- the trees are very unequal here, unlike the Milky-Way forest of r02z
OUT=gpurun_out
mkdir -p $OUT
for kn in "GLC_STREAM_PRIORITY_EXPRESS=0" "GLC_STREAM_PRIORITY_EXPRESS=100" "GLC_STREAM_PRIORITY_EXPRESS=300" "GLC_STREAM_PRIORITY_EXPRESS=30"; do
  env $kn FOREST_KIND=volume GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 12500 0 > $OUT/r2av_forest.log 2>&1; echo "$kn exit $?"
  grep "FOREST\|forest async" $OUT/r2av_forest.log | tail -3 | cut -c1-200
done
