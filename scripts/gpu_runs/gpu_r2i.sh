#!/bin/bash
# round-2 GPU call I: stream test with the slice log (finish after lane passes over-counts), root-steps variants on the
# 10^6-node pass, ncu --set full capture of a bulk machine slice (report kept for source-level attribution)
OUT=gpurun_out
mkdir -p $OUT
echo "== I1 stream test, adaptive ticks, slice log"
GLC_SLICE_LOG=1 timeout 300 python -m pytest "tests/test_gpu_stream.py::test_stream_equals_batch" -m gpu -x -q > $OUT/r2i_stream.log 2>&1; echo "exit $?"
grep "glc stream\|glc slice\|glc drain\|passed\|failed" $OUT/r2i_stream.log | cut -c1-220 | tail -60
echo "== I2 root-steps variants"
for v in base rs2 rs4; do
  LIB=galacticus_b200/libglcb200_$v.so; [ $v = base ] && LIB=galacticus_b200/libglcb200.so
  timeout 300 python scripts/knobs.py 1000000 LIB=$LIB GLC_SLICE_LOG=1 2> $OUT/r2i_knobs_$v.err | grep KNOBS
  grep "glc slice\|glc drain" $OUT/r2i_knobs_$v.err | tail -40 | head -12 | cut -c1-200
done
echo "== I3 ncu machine_kernel bulk slice"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:machine_kernel -s 1 -c 1 -f -o $OUT/r02i_machine \
  python bench.py --steps 1 --warmup 0 --nodes 1000000 --trees 0 --volume-trees 0 --cpu-sample 1000 > /dev/null 2> $OUT/r2i_ncu_machine_err.log
echo "exit $?"
ls -la $OUT | tail -8
