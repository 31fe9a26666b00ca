#!/bin/bash
# round-2 GPU call P: explicit fma() in glc_detmath.h (both sides) -- full GPU suite, 10^6-node pass, forest 1000
OUT=gpurun_out
mkdir -p $OUT
echo "== P1 full GPU suite"
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > $OUT/r2p_pytest_gpu.log 2>&1; echo "exit $?"; tail -10 $OUT/r2p_pytest_gpu.log
echo "== P2 10^6-node pass"
timeout 300 python scripts/knobs.py 1000000 GLC_SLICE_LOG=1 2> $OUT/r2p_knobs.err | grep KNOBS
grep "glc slice\|glc drain" $OUT/r2p_knobs.err | tail -18 | cut -c1-170
echo "== P3 forest 1000"
GLC_FOREST_LOG=1 timeout 400 python scripts/forest_bench.py 1000 256 > $OUT/r2p_mw1000.log 2>&1; echo "exit $?"
grep "FOREST\|forest async" $OUT/r2p_mw1000.log | tail -3 | cut -c1-400
grep -o '"cpu": {[^}]*}' $OUT/r2p_mw1000.log
