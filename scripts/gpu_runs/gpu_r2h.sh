#!/bin/bash
# round-2 GPU call H (after the container was re-created): full GPU suite, asynchronous vs bulk-synchronous forest
# schedule on configs[1] (1000 MW trees), configs[3] (12500 volume trees) and 4000 MW trees, then the default bench lines
OUT=gpurun_out
mkdir -p $OUT
echo "== H1 full GPU suite"
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/r2h_pytest_gpu.log 2>&1; echo "exit $?"; tail -14 $OUT/r2h_pytest_gpu.log
for mode in 1 0; do
  echo "== H2 1000 MW trees, GLC_FOREST_ASYNC=$mode"
  GLC_FOREST_ASYNC=$mode GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2h_mw1000_async$mode.log 2>&1; echo "exit $?"
  grep "FOREST\|forest async" $OUT/r2h_mw1000_async$mode.log | cut -c1-400
done
for mode in 1 0; do
  echo "== H3 volume 12500 trees, GLC_FOREST_ASYNC=$mode"
  FOREST_KIND=volume GLC_FOREST_ASYNC=$mode GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 12500 0 > $OUT/r2h_vol12500_async$mode.log 2>&1; echo "exit $?"
  grep "FOREST\|forest async" $OUT/r2h_vol12500_async$mode.log | cut -c1-400
done
echo "== H4 4000 MW trees, async"
GLC_FOREST_ASYNC=1 GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 4000 0 > $OUT/r2h_mw4000_async1.log 2>&1; echo "exit $?"
grep "FOREST\|forest async" $OUT/r2h_mw4000_async1.log | cut -c1-400
echo "== H5 bench"
timeout 900 python bench.py > $OUT/r2h_bench_line.json 2> $OUT/r2h_bench_err.log; echo "exit $?"; cut -c1-600 $OUT/r2h_bench_line.json
