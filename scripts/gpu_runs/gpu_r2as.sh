#!/bin/bash
# round-2 GPU call AS: a block of a dense drain pass drops the block vote once at most N nodes are left in it
OUT=gpurun_out
mkdir -p $OUT
for kn in 8 0 16 32 64 8 0; do
  timeout 300 python scripts/knobs.py 1000000 GLC_DRAIN_SYNC_MIN=$kn | grep KNOBS
done
for kn in 8 32 0; do
  GLC_DRAIN_SYNC_MIN=$kn GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 0 > $OUT/r2as_forest.log 2>&1; echo "sync min $kn exit $?"
  grep "FOREST\|forest async" $OUT/r2as_forest.log | tail -3 | cut -c1-200
done
timeout 900 python -m pytest tests/test_gpu_machine_scale.py tests/test_gpu_standard.py tests/test_gpu_stream.py -m gpu -x -q -k "not forest_4000" > $OUT/r2as_pytest.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/r2as_pytest.log
