#!/bin/bash
# round-2 GPU call AW: lane passes spread over the resident warps (stream_spread 1 / 2), volume forest and Milky-Way forest, with the
# cooperative GK15 pass (which favours thin warps)
OUT=gpurun_out
mkdir -p $OUT
for kn in "GLC_STREAM_SPREAD=1" "GLC_STREAM_SPREAD=2" "GLC_STREAM_SPREAD=0"; do
  env $kn FOREST_KIND=volume GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 12500 0 > $OUT/r2aw_forest.log 2>&1; echo "volume $kn exit $?"
  grep "FOREST\|forest async" $OUT/r2aw_forest.log | tail -3 | cut -c1-200
done
for kn in "GLC_STREAM_SPREAD=2"; do
  env $kn GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2aw_forest.log 2>&1; echo "milky way $kn exit $?"
  grep "FOREST\|forest async" $OUT/r2aw_forest.log | tail -3 | cut -c1-200
done
