#!/bin/bash
# round-2 GPU call M: express warps in the streaming lane passes (oldest / longest nodes alone in a warp)
OUT=gpurun_out
mkdir -p $OUT
echo "== M0 parity: forest + stream tests"
timeout 600 python -m pytest tests/test_forest.py tests/test_gpu_stream.py -m gpu -x -q > $OUT/r2m_pytest.log 2>&1; echo "exit $?"; tail -3 $OUT/r2m_pytest.log
for cfg in "24 48" "0 48" "8 48" "24 96" "64 48"; do
  set -- $cfg
  GLC_STREAM_EXPRESS=$1 GLC_STREAM_EXPRESS_BUDGET=$2 GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2m_mw1000_$1_$2.log 2>&1; echo "express=$1 budget=$2 exit $?"
  grep "FOREST\|forest async" $OUT/r2m_mw1000_$1_$2.log | tail -3 | cut -c1-250
done
