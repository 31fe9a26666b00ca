#!/bin/bash
# round-2 GPU call J: stream tests after the stale-slot fix; 10^6-node pass with the unit word in a register + staged pow table
# (base), 384/512-thread machine blocks, 2 Brent steps per unit; forest 1000 async under three machine/lane thresholds
OUT=gpurun_out
mkdir -p $OUT
echo "== J1 stream tests"
timeout 300 python -m pytest tests/test_gpu_stream.py -m gpu -x -q > $OUT/r2j_stream.log 2>&1; echo "exit $?"; tail -3 $OUT/r2j_stream.log
echo "== J2 variants on the 10^6-node pass"
for v in base rs2 t384 t512; do
  LIB=galacticus_b200/libglcb200_$v.so; [ $v = base ] && LIB=galacticus_b200/libglcb200.so
  timeout 300 python scripts/knobs.py 1000000 LIB=$LIB GLC_SLICE_LOG=1 2> $OUT/r2j_knobs_$v.err | grep KNOBS
  grep "glc slice" $OUT/r2j_knobs_$v.err | tail -14 | head -8 | cut -c1-170
done
echo "== J3 forest 1000 async: machine/lane threshold"
for th in 120000 40000 12000; do
  GLC_STREAM_MACHINE_ABOVE=$th GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2j_mw1000_th$th.log 2>&1; echo "th=$th exit $?"
  grep "FOREST\|forest async" $OUT/r2j_mw1000_th$th.log | cut -c1-260
done
