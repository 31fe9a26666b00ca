#!/bin/bash
# round-2 GPU call AD: one copy of the Brent state machine per kernel (GLC_NOINLINE_BRENT) -- the dense drain pass is bound by
# instruction fetch (no_instruction 12.8 stall cycles per issue, profiles/r02ac): does a smaller kernel help?
OUT=gpurun_out
mkdir -p $OUT
for lib in galacticus_b200/libglcb200_nib.so galacticus_b200/libglcb200.so; do
  timeout 300 python scripts/knobs.py 1000000 LIB=$lib GLC_SLICE_LOG=1 2> $OUT/r2ad_knobs.err | grep KNOBS
  grep "(hold)\|glc drain" $OUT/r2ad_knobs.err | tail -3 | cut -c1-150
  GLC_LIB_PATH=$PWD/$lib GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2ad_forest.log 2>&1; echo "$lib exit $?"
  grep "FOREST\|forest async" $OUT/r2ad_forest.log | tail -2 | cut -c1-200
done
