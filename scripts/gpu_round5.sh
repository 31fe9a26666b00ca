#!/bin/bash
# last GPU call of the round: parity after the scheduler change, bench lines of the final build, 4000-tree forest
TAG=${1:-r01k}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 300 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest_gpu.log
echo "== bench"; timeout 400 python bench.py > $OUT/${TAG}_bench_line.json 2> $OUT/${TAG}_bench_err.log; echo "bench exit $?"; cut -c1-200 $OUT/${TAG}_bench_line.json; tail -3 $OUT/${TAG}_bench_err.log
echo "== forest 4000"; GLC_FOREST_LOG=1 timeout 170 python scripts/forest_bench.py 4000 0 > $OUT/${TAG}_forest_4000.log 2>&1; echo "forest exit $?"; grep FOREST $OUT/${TAG}_forest_4000.log | cut -c1-400; grep -c "glc forest" $OUT/${TAG}_forest_4000.log; tail -2 $OUT/${TAG}_forest_4000.log | cut -c1-300
