#!/bin/bash
# round-2 GPU call Y: 64-thread blocks for drain_kernel / evolve_kernel (2 warps per block: the nodes of a lane pass land on
# twice as many SMs with half as many warps each)
OUT=gpurun_out
mkdir -p $OUT
for lib in galacticus_b200/libglcb200_b64.so galacticus_b200/libglcb200.so; do
  GLC_LIB_PATH=$PWD/$lib GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2y_forest.log 2>&1; echo "$lib exit $?"
  grep "FOREST\|forest async" $OUT/r2y_forest.log | tail -3 | cut -c1-230
  timeout 300 python scripts/knobs.py 1000000 LIB=$lib | grep KNOBS
done
