#!/bin/bash
# round-2 GPU call BC: hand-over with and without the express launch, alternating (run-to-run scatter is +-40 ms)
for i in 1 2 3; do
for kn in "GLC_DRAIN_EXPRESS=0" "GLC_DRAIN_EXPRESS=1"; do
  timeout 300 python scripts/knobs.py 1000000 $kn | grep KNOBS
done
done
