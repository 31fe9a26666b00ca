#!/bin/bash
# round-2 GPU call Q: hand-over threshold x internal slice budget of the hybrid pass (10^6 nodes)
OUT=gpurun_out
mkdir -p $OUT
for cfg in "GLC_DRAIN_BELOW=120000 GLC_HYBRID_BUDGET=4096" "GLC_DRAIN_BELOW=60000 GLC_HYBRID_BUDGET=1024" "GLC_DRAIN_BELOW=30000 GLC_HYBRID_BUDGET=1024" "GLC_DRAIN_BELOW=15000 GLC_HYBRID_BUDGET=512" "GLC_DRAIN_BELOW=120000 GLC_HYBRID_BUDGET=1024"; do
  tag=$(echo $cfg | tr ' =' '__')
  timeout 300 python scripts/knobs.py 1000000 $cfg GLC_SLICE_LOG=1 2> $OUT/r2q_$tag.err | grep KNOBS
  grep "(hold)\|glc drain" $OUT/r2q_$tag.err | tail -4 | cut -c1-170
done
