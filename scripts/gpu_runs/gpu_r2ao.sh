#!/bin/bash
# round-2 GPU call AO: early express with an age condition on the prediction, shorter machine slices
OUT=gpurun_out
mkdir -p $OUT
for kn in "GLC_EARLY_EXPRESS_AGE=24" "GLC_EARLY_EXPRESS_AGE=24 GLC_HYBRID_BUDGET=1024" "GLC_EARLY_EXPRESS_AGE=40 GLC_EARLY_EXPRESS_STEPS=100 GLC_HYBRID_BUDGET=1024" "GLC_EARLY_EXPRESS_STEPS=0 GLC_HYBRID_BUDGET=1024" "GLC_EARLY_EXPRESS_AGE=12 GLC_EARLY_EXPRESS_STEPS=150 GLC_HYBRID_BUDGET=2048"; do
  timeout 300 python scripts/knobs.py 1000000 $kn GLC_SLICE_LOG=1 2> $OUT/r2ao_knobs.err | grep KNOBS
  grep "(hold)\|early express\] [0-9]\|dense part" $OUT/r2ao_knobs.err | tail -3 | cut -c1-160
  grep "early express\] launch" $OUT/r2ao_knobs.err | tail -12 | cut -c1-100 | tr '\n' ';'; echo
done
