#!/bin/bash
# round-2 GPU call Z: express warps chosen by the tree scheduler's priority (halo mass: the main branches) in the lane passes
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_forest.py -m gpu -x -q > $OUT/r2z_pytest.log 2>&1; echo "pytest (default) exit $?"
GLC_STREAM_PRIORITY_EXPRESS=592 timeout 600 python -m pytest tests/test_forest.py -m gpu -x -q > $OUT/r2z_pytest_prio.log 2>&1; echo "pytest (priority express) exit $?"; tail -2 $OUT/r2z_pytest_prio.log
for cfg in "592 24" "592 48" "1000 24" "200 24" "0 24"; do
  set -- $cfg
  GLC_STREAM_PRIORITY_EXPRESS=$1 GLC_STREAM_EXPRESS_BUDGET=$2 GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2z_mw1000_$1_$2.log 2>&1; echo "priority express=$1 budget=$2 exit $?"
  grep "FOREST\|forest async" $OUT/r2z_mw1000_$1_$2.log | tail -3 | cut -c1-250
done
