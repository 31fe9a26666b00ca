#!/bin/bash
# round-2 GPU call C: the bounded MPMC queue protocol under the ledger (4000-tree forest), then parity + timing of the product build
OUT=gpurun_out
mkdir -p $OUT
LED=$PWD/galacticus_b200/libglcb200_ledger.so
echo "== C1 ledger build, 4000 trees"
GLC_LIB_PATH=$LED GLC_LEDGER_DUMP=$OUT/r2c_stuck_slots.bin GLC_SLICE_LOG=0 GLC_FOREST_LOG=1 \
  timeout 400 python scripts/forest_bench.py 4000 0 > $OUT/r2c_c1.log 2>&1
echo "exit $?"; grep -c "glc forest" $OUT/r2c_c1.log; grep -n "ledger\|held by\|never fetched\|FOREST\|Error\|failed\|Warning" $OUT/r2c_c1.log | cut -c1-420 | head -40
echo "== C2 product build: pytest -m gpu"
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/r2c_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/r2c_pytest_gpu.log
echo "== C3 product build: 4000 trees"
GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 4000 0 > $OUT/r2c_c3.log 2>&1
echo "exit $?"; grep -n "FOREST\|Error\|failed\|Warning" $OUT/r2c_c3.log | cut -c1-600 | head
echo "== C4 product build: bench (no trees)"
timeout 500 python bench.py --trees 0 > $OUT/r2c_bench_line.json 2> $OUT/r2c_bench_err.log; echo "bench exit $?"; cut -c1-300 $OUT/r2c_bench_line.json; tail -3 $OUT/r2c_bench_err.log
