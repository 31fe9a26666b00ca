#!/bin/bash
# round-2 GPU call K: nodes spread over all resident warps in the drain / lane passes (drainLanes); staged pow table on/off x
# 1 or 2 Brent steps per unit on the 10^6-node pass; forest 1000 async with and without spreading
OUT=gpurun_out
mkdir -p $OUT
echo "== K1 variants on the 10^6-node pass"
for v in base ns rs2 nsrs2; do
  LIB=galacticus_b200/libglcb200_$v.so; [ $v = base ] && LIB=galacticus_b200/libglcb200.so
  timeout 300 python scripts/knobs.py 1000000 LIB=$LIB GLC_SLICE_LOG=1 2> $OUT/r2k_knobs_$v.err | grep KNOBS
  grep "glc drain" $OUT/r2k_knobs_$v.err | tail -6 | cut -c1-170
done
echo "== K1b base, no spreading"
timeout 300 python scripts/knobs.py 1000000 GLC_DRAIN_SPREAD=0 GLC_SLICE_LOG=1 2> $OUT/r2k_knobs_nospread.err | grep KNOBS
echo "== K2 forest 1000 async: spread on / off"
for sp in 1 0; do
  GLC_DRAIN_SPREAD=$sp GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2k_mw1000_spread$sp.log 2>&1; echo "spread=$sp exit $?"
  grep "FOREST\|forest async" $OUT/r2k_mw1000_spread$sp.log | cut -c1-260
done
echo "== K3 GPU forest + stream + machine tests"
timeout 900 python -m pytest tests/test_forest.py tests/test_gpu_stream.py tests/test_gpu_standard.py -m gpu -x -q > $OUT/r2k_pytest.log 2>&1; echo "exit $?"; tail -3 $OUT/r2k_pytest.log
