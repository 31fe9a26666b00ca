#!/bin/bash
# round-2 GPU call AZ: final rule for lane passes (one block per SM from 593 nodes on): streaming and forest parity tests, forests
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_forest.py tests/test_gpu_standard.py -m gpu -x -q > $OUT/r2az_pytest.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/r2az_pytest.log
FOREST_KIND=volume GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 12500 0 > $OUT/r2az_forest.log 2>&1; echo "volume exit $?"
grep "FOREST\|forest async" $OUT/r2az_forest.log | tail -3 | cut -c1-200
GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2az_forest.log 2>&1; echo "milky way exit $?"
grep "FOREST\|forest async" $OUT/r2az_forest.log | tail -3 | cut -c1-200
