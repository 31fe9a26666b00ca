#!/bin/bash
# round-2 GPU call U: new GPU tests (error report, general beta) + the full suite + the 10^6-node pass of the same build
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r2u_pytest_gpu.log 2>&1; echo "exit $?"; tail -4 $OUT/r2u_pytest_gpu.log
timeout 300 python scripts/knobs.py 1000000 | grep KNOBS
