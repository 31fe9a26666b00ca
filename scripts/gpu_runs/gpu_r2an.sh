#!/bin/bash
# round-2 GPU call AN: early express -- the predicted-longest nodes leave the machine between its slices and are finished one per
# warp beside the following slices
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_machine_scale.py -m gpu -x -q -k "early or 600k" > $OUT/r2an_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r2an_pytest.log
for kn in "GLC_EARLY_EXPRESS_STEPS=60" "GLC_EARLY_EXPRESS_STEPS=0" "GLC_EARLY_EXPRESS_STEPS=40" "GLC_EARLY_EXPRESS_STEPS=100" "GLC_EARLY_EXPRESS_STEPS=60 GLC_EARLY_EXPRESS_SMEM_KB=0"; do
  timeout 300 python scripts/knobs.py 1000000 $kn GLC_SLICE_LOG=1 2> $OUT/r2an_knobs.err | grep KNOBS
  grep "(hold)\|early express\] [0-9]\|dense part" $OUT/r2an_knobs.err | tail -3 | cut -c1-160
done
