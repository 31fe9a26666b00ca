#!/bin/bash
# round-2 GPU call N: lane-pass lists sorted by kind of node; spread over one block per SM; forest size scan
OUT=gpurun_out
mkdir -p $OUT
for cfg in "1 0" "0 0" "1 2" "0 2"; do
  set -- $cfg
  GLC_STREAM_SORT=$1 GLC_STREAM_SPREAD=$2 GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2n_mw1000_sort$1_spread$2.log 2>&1; echo "sort=$1 spread=$2 exit $?"
  grep "FOREST\|forest async" $OUT/r2n_mw1000_sort$1_spread$2.log | tail -3 | cut -c1-250
done
for nt in 60 250; do
  GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py $nt 0 > $OUT/r2n_mw$nt.log 2>&1; echo "trees=$nt exit $?"
  grep "FOREST\|forest async" $OUT/r2n_mw$nt.log | tail -3 | cut -c1-250
done
timeout 300 python -m pytest tests/test_forest.py tests/test_gpu_stream.py -m gpu -x -q > $OUT/r2n_pytest.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/r2n_pytest.log
