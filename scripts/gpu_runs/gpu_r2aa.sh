#!/bin/bash
# round-2 GPU call AA: last validation of the committed build -- full GPU suite, smoke, 4000-tree forest
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r2aa_pytest_gpu.log 2>&1; echo "exit $?"; tail -3 $OUT/r2aa_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2aa_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/r2aa_smoke.log
GLC_FOREST_LOG=1 timeout 600 python scripts/forest_bench.py 4000 0 > $OUT/r2aa_mw4000.log 2>&1; echo "exit $?"
grep "FOREST\|forest async" $OUT/r2aa_mw4000.log | tail -3 | cut -c1-330
