#!/bin/bash
# round-2 GPU call AJ: knobs again now that a GK15 pass is shared by 15 lanes (lone lanes are relatively faster), then an ncu capture
# of the dense drain pass of the new build
OUT=gpurun_out
mkdir -p $OUT
for kn in "GLC_DRAIN_BELOW=200000" "GLC_DRAIN_BELOW=300000" "GLC_DRAIN_LANES_MAX=16" "GLC_DRAIN_BELOW=80000"; do
  timeout 300 python scripts/knobs.py 1000000 $kn GLC_SLICE_LOG=1 2> $OUT/r2aj_knobs.err | grep KNOBS
  grep "(hold)" $OUT/r2aj_knobs.err | tail -1 | cut -c1-120
done
for kn in "GLC_STREAM_SPREAD=1" "GLC_STREAM_DENSE_BUDGET=24" "GLC_STREAM_DENSE_BUDGET=6" "GLC_STREAM_EXPRESS=400"; do
  env $kn GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2aj_forest.log 2>&1; echo "$kn exit $?"
  grep "FOREST\|forest async" $OUT/r2aj_forest.log | tail -3 | cut -c1-200
done
GLC_DRAIN_EXPRESS=0 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:drain_kernel -s 0 -c 1 -f -o $OUT/r02aj_drain_dense \
  python bench.py --steps 1 --warmup 0 --nodes 1000000 --trees 0 --volume-trees 0 --cpu-sample 1000 > /dev/null 2> $OUT/r2aj_ncu_err.log
echo "ncu exit $?"; ls -la $OUT | grep r02aj
