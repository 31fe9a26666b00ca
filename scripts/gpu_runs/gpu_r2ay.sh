#!/bin/bash
# round-2 GPU call AY: one block per SM also between 592 and 1184 nodes (stream_spread 3); hand-over without the express launch
OUT=gpurun_out
mkdir -p $OUT
for kn in "GLC_STREAM_SPREAD=3" "GLC_STREAM_SPREAD=2"; do
  env $kn FOREST_KIND=volume GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 12500 0 > $OUT/r2ay_forest.log 2>&1; echo "volume $kn exit $?"
  grep "FOREST\|forest async" $OUT/r2ay_forest.log | tail -3 | cut -c1-200
done
for kn in "GLC_STREAM_SPREAD=3"; do
  env $kn GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2ay_forest.log 2>&1; echo "milky way $kn exit $?"
  grep "FOREST\|forest async" $OUT/r2ay_forest.log | tail -3 | cut -c1-200
done
for kn in "GLC_DRAIN_EXPRESS=0" "GLC_DRAIN_EXPRESS=0 GLC_DRAIN_SPREAD=0" "GLC_DRAIN_EXPRESS=1"; do
  timeout 300 python scripts/knobs.py 1000000 $kn | grep KNOBS
done
