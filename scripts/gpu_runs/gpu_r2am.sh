#!/bin/bash
# round-2 GPU call AM: build with the block vote (level 2) and GK15 passes across lanes -- full GPU suite, both bench arms, launch list,
# the 4000-tree forest
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r2am_pytest_gpu.log 2>&1; echo "exit $?"; tail -3 $OUT/r2am_pytest_gpu.log
timeout 1200 python bench.py > $OUT/r2am_bench_line.json 2> $OUT/r2am_bench_err.log; echo "exit $?"; cut -c1-260 $OUT/r2am_bench_line.json
timeout 1200 python bench.py --impl reference > $OUT/r2am_bench_reference_line.json 2> $OUT/r2am_bench_ref_err.log; echo "exit $?"; cut -c1-200 $OUT/r2am_bench_reference_line.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/r02am_launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --nodes 300000 --trees 0 --volume-trees 0 --cpu-sample 1000 > $OUT/r2am_launch_bench.json 2> $OUT/r2am_launch_err.log
echo "exit $?"; wc -l $OUT/r02am_launches_bench.csv
GLC_FOREST_LOG=1 timeout 600 python scripts/forest_bench.py 4000 0 > $OUT/r2am_forest4000.log 2>&1; echo "exit $?"
grep "FOREST\|forest async" $OUT/r2am_forest4000.log | tail -3 | cut -c1-220
