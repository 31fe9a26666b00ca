#!/bin/bash
# round-2 GPU call V: unbounded dense drain pass (stragglers end up alone in their warps once their warp-mates are done)
OUT=gpurun_out
mkdir -p $OUT
for cfg in "GLC_DRAIN_DENSE_BUDGET=1000000" "GLC_DRAIN_DENSE_BUDGET=1000000 GLC_DRAIN_EXPRESS=0" "GLC_DRAIN_DENSE_BUDGET=4096" "GLC_DRAIN_EXPRESS=0"; do
  tag=$(echo $cfg | tr ' =' '__')
  timeout 300 python scripts/knobs.py 1000000 $cfg GLC_SLICE_LOG=1 2> $OUT/r2v_$tag.err | grep KNOBS
  grep "glc drain" $OUT/r2v_$tag.err | tail -3 | cut -c1-170
done
