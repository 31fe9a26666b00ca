#!/bin/bash
# round-2 GPU call AP: dense part of the hand-over list sorted by predicted steps (longest first, dealt out across the warps)
OUT=gpurun_out
mkdir -p $OUT
for kn in "GLC_DRAIN_SORT_DENSE=1" "GLC_DRAIN_SORT_DENSE=0" "GLC_DRAIN_SORT_DENSE=1 GLC_DRAIN_AGE_WEIGHT=1" "GLC_DRAIN_SORT_DENSE=1"; do
  timeout 300 python scripts/knobs.py 1000000 $kn GLC_SLICE_LOG=1 2> $OUT/r2ap_knobs.err | grep KNOBS
  grep "(hold)\|dense part" $OUT/r2ap_knobs.err | tail -2 | cut -c1-160
done
timeout 900 python -m pytest tests/test_gpu_machine_scale.py tests/test_gpu_standard.py -m gpu -x -q -k "not forest" > $OUT/r2ap_pytest.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/r2ap_pytest.log
