#!/bin/bash
# round-2 GPU call A: baseline parity + the lost-node reproducer under the ownership ledger (debug build)
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/r2a_smi.txt 2>&1
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q > $OUT/r2a_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/r2a_pytest_gpu.log
echo "== forest 4000 under the ledger"
GLC_LIB_PATH=$PWD/galacticus_b200/libglcb200_ledger.so GLC_SLICE_LOG=1 GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 4000 0 > $OUT/r2a_forest_4000_ledger.log 2>&1
echo "forest exit $?"; grep -n "ledger\|held by\|FOREST\|Error\|failed" $OUT/r2a_forest_4000_ledger.log | head -60; tail -3 $OUT/r2a_forest_4000_ledger.log | cut -c1-300
