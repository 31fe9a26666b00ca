#!/bin/bash
# round-2 GPU call AR: subsets of the phase barriers of the block vote (levels 8 + mask: 1 after the structure solve, 2 after star
# formation, 4 after the cooling radius), with the cooperative GK15 pass in place
OUT=gpurun_out
mkdir -p $OUT
for bs in 2 9 13 1 11 12 2; do
  timeout 300 python scripts/knobs.py 1000000 GLC_DRAIN_BLOCK_SYNC=$bs | grep KNOBS
done
for bs in 9 2; do
  GLC_DRAIN_BLOCK_SYNC=$bs GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2ar_forest.log 2>&1; echo "level $bs exit $?"
  grep "FOREST\|forest async" $OUT/r2ar_forest.log | tail -3 | cut -c1-200
done
