#!/bin/bash
# round-2 GPU call AI: the 15 points of a GK15 pass across lanes (qag15_coop) vs the sequential pass (-DGLC_NO_COOP_QAG), parity
# tests first; 168-register build (3 blocks per SM) for comparison
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_standard.py tests/test_gpu_stream.py tests/test_forest.py -m gpu -x -q > $OUT/r2ai_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r2ai_pytest.log
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 300 python scripts/knobs.py 1000000 GLC_SLICE_LOG=1 2> $OUT/r2ai_knobs.err | grep KNOBS
  grep "(hold)" $OUT/r2ai_knobs.err | tail -1 | cut -c1-120
  env "$@" GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2ai_forest.log 2>&1; echo "$label exit $?"
  grep "FOREST\|forest async" $OUT/r2ai_forest.log | tail -3 | cut -c1-200
}
run "coop" GLC_X=1
run "sequential" GLC_LIB_PATH=$PWD/galacticus_b200/libglcb200_nocoop.so
run "coop, 3 blocks per SM" GLC_LIB_PATH=$PWD/galacticus_b200/libglcb200_mb3.so
