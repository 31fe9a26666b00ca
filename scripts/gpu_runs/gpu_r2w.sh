#!/bin/bash
# round-2 GPU call W: unbounded dense drain pass without the express kernel, 16 / 8 nodes per warp on all resident warps
OUT=gpurun_out
mkdir -p $OUT
for cfg in "GLC_DRAIN_DENSE_BUDGET=1000000" "GLC_DRAIN_DENSE_BUDGET=1000000 GLC_DRAIN_EXPRESS=0 GLC_DRAIN_LANES_MAX=16" "GLC_DRAIN_DENSE_BUDGET=1000000 GLC_DRAIN_EXPRESS=0 GLC_DRAIN_LANES_MAX=8" "GLC_DRAIN_DENSE_BUDGET=1000000 GLC_DRAIN_EXPRESS=0 GLC_DRAIN_LANES_MAX=4" "GLC_DRAIN_DENSE_BUDGET=1000000 GLC_DRAIN_BELOW=200000"; do
  tag=$(echo $cfg | tr ' =' '__')
  timeout 300 python scripts/knobs.py 1000000 $cfg GLC_SLICE_LOG=1 2> $OUT/r2w_$tag.err | grep KNOBS
  grep "glc drain" $OUT/r2w_$tag.err | tail -2 | cut -c1-170
done
