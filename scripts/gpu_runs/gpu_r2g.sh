#!/bin/bash
# round-2 GPU call G: asynchronous forest schedule -- parity, then timing against the bulk-synchronous rounds
OUT=gpurun_out
mkdir -p $OUT
echo "== G1 forest GPU tests"
timeout 600 python -m pytest tests/test_forest.py -m gpu -x -q > $OUT/r2g_pytest_forest.log 2>&1; echo "exit $?"; tail -4 $OUT/r2g_pytest_forest.log
for mode in 1 0; do
  echo "== G2 1000 MW trees, GLC_FOREST_ASYNC=$mode"
  GLC_FOREST_ASYNC=$mode GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2g_mw1000_async$mode.log 2>&1; echo "exit $?"
  grep "FOREST\|forest async" $OUT/r2g_mw1000_async$mode.log | cut -c1-330
done
for mode in 1 0; do
  echo "== G3 volume 12500 trees, GLC_FOREST_ASYNC=$mode"
  FOREST_KIND=volume GLC_FOREST_ASYNC=$mode GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 12500 0 > $OUT/r2g_vol12500_async$mode.log 2>&1; echo "exit $?"
  grep "FOREST\|forest async" $OUT/r2g_vol12500_async$mode.log | cut -c1-330
done
echo "== G4 4000 MW trees, async"
GLC_FOREST_ASYNC=1 GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 4000 0 > $OUT/r2g_mw4000_async1.log 2>&1; echo "exit $?"
grep "FOREST\|forest async" $OUT/r2g_mw4000_async1.log | cut -c1-330
