#!/bin/bash
# round-2 GPU call AE: the warps of a drain block start every evaluation together (GLC_DRAIN_BLOCK_SYNC=1): shared instruction fetch
OUT=gpurun_out
mkdir -p $OUT
for bs in 1 0; do
  timeout 300 python scripts/knobs.py 1000000 GLC_DRAIN_BLOCK_SYNC=$bs GLC_SLICE_LOG=1 2> $OUT/r2ae_knobs.err | grep KNOBS
  grep "(hold)\|glc drain" $OUT/r2ae_knobs.err | tail -3 | cut -c1-150
  GLC_DRAIN_BLOCK_SYNC=$bs GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2ae_forest.log 2>&1; echo "block sync $bs exit $?"
  grep "FOREST\|forest async" $OUT/r2ae_forest.log | tail -2 | cut -c1-200
done
GLC_DRAIN_BLOCK_SYNC=1 timeout 600 python -m pytest tests/test_forest.py tests/test_gpu_stream.py tests/test_gpu_standard.py -m gpu -x -q > $OUT/r2ae_pytest.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/r2ae_pytest.log
