#!/bin/bash
# round-2 GPU call AK: evaluations per lane in a streaming lane pass (stream_dense_budget, default 12) with the cooperative GK15 build
OUT=gpurun_out
mkdir -p $OUT
for kn in "GLC_STREAM_DENSE_BUDGET=12" "GLC_STREAM_DENSE_BUDGET=24" "GLC_STREAM_DENSE_BUDGET=48" "GLC_STREAM_DENSE_BUDGET=24 GLC_STREAM_SPARSE_BUDGET=64" "GLC_STREAM_DENSE_BUDGET=18"; do
  env $kn GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2ak_forest.log 2>&1; echo "$kn exit $?"
  grep "FOREST\|forest async" $OUT/r2ak_forest.log | tail -3 | cut -c1-200
done
