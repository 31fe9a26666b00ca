#!/bin/bash
# round-2 GPU call O: the tail of the 10^6-node pass -- dense budget and age-weighted express selection
OUT=gpurun_out
mkdir -p $OUT
for cfg in "GLC_DRAIN_DENSE_BUDGET=1024" "GLC_DRAIN_DENSE_BUDGET=256" "GLC_DRAIN_DENSE_BUDGET=96" "GLC_DRAIN_AGE_WEIGHT=1" "GLC_DRAIN_AGE_WEIGHT=6" "GLC_DRAIN_AGE_WEIGHT=6 GLC_DRAIN_DENSE_BUDGET=256"; do
  tag=$(echo $cfg | tr ' =' '__')
  timeout 300 python scripts/knobs.py 1000000 $cfg GLC_SLICE_LOG=1 2> $OUT/r2o_$tag.err | grep KNOBS
  grep "glc drain" $OUT/r2o_$tag.err | tail -5 | cut -c1-170
done
