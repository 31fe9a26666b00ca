#!/bin/bash
# round-2 GPU call BF: launch list of the final build (no express launch at the hand-over)
OUT=gpurun_out
mkdir -p $OUT
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/r02bf_launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --nodes 300000 --trees 0 --volume-trees 0 --cpu-sample 1000 > $OUT/r2bf_launch_bench.json 2> $OUT/r2bf_launch_err.log
echo "exit $?"; wc -l $OUT/r02bf_launches_bench.csv
