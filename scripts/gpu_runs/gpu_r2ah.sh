#!/bin/bash
# round-2 GPU call AH: vote level 3 (structure-solve passes vote per block), and 256-thread blocks (one block per SM: all 8
# warps of the SM in step) at levels 2 and 3
OUT=gpurun_out
mkdir -p $OUT
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 300 python scripts/knobs.py 1000000 GLC_SLICE_LOG=1 2> $OUT/r2ah_knobs.err | grep KNOBS
  grep "(hold)" $OUT/r2ah_knobs.err | tail -1 | cut -c1-120
  env "$@" GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2ah_forest.log 2>&1; echo "$label exit $?"
  grep "FOREST\|forest async" $OUT/r2ah_forest.log | tail -2 | cut -c1-200
}
run "128x2 level 3" GLC_DRAIN_BLOCK_SYNC=3
run "256x1 level 2" GLC_DRAIN_BLOCK_SYNC=2 GLC_LIB_PATH=$PWD/galacticus_b200/libglcb200_b256.so
run "256x1 level 3" GLC_DRAIN_BLOCK_SYNC=3 GLC_LIB_PATH=$PWD/galacticus_b200/libglcb200_b256.so
run "128x2 level 2" GLC_DRAIN_BLOCK_SYNC=2
