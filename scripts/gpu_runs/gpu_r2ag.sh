#!/bin/bash
# round-2 GPU call AG: vote levels of the drain kernel (GLC_DRAIN_BLOCK_SYNC): 1 = block vote at the loop exit, 2 = + block
# barriers between the phases of the rate function, 3 = + the data-dependent loops vote per block
OUT=gpurun_out
mkdir -p $OUT
for bs in 3 2 1; do
  timeout 300 python scripts/knobs.py 1000000 GLC_DRAIN_BLOCK_SYNC=$bs GLC_SLICE_LOG=1 2> $OUT/r2ag_knobs.err | grep KNOBS
  grep "(hold)\|glc drain" $OUT/r2ag_knobs.err | tail -3 | cut -c1-150
  GLC_DRAIN_BLOCK_SYNC=$bs GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2ag_forest.log 2>&1; echo "block sync $bs exit $?"
  grep "FOREST\|forest async" $OUT/r2ag_forest.log | tail -2 | cut -c1-200
done
for bs in 3 2; do
GLC_DRAIN_BLOCK_SYNC=$bs timeout 600 python -m pytest tests/test_forest.py tests/test_gpu_stream.py tests/test_gpu_standard.py -m gpu -x -q > $OUT/r2ag_pytest$bs.log 2>&1; echo "pytest level $bs exit $?"; tail -2 $OUT/r2ag_pytest$bs.log
done
