#!/bin/bash
# round-2 GPU call R: ncu captures of the final build -- last (one node per warp) drain pass, a bulk machine slice, launch list
OUT=gpurun_out
mkdir -p $OUT
echo "== R1 drain_kernel, last pass (launch index 2 of the pass: express, dense, sparse)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:drain_kernel -s 2 -c 1 -f -o $OUT/r02r_drain_sparse \
  python bench.py --steps 1 --warmup 0 --nodes 1000000 --trees 0 --volume-trees 0 --cpu-sample 1000 > /dev/null 2> $OUT/r2r_ncu_drain_err.log
echo "exit $?"
echo "== R2 machine_kernel, second slice"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:machine_kernel -s 1 -c 1 -f -o $OUT/r02r_machine \
  python bench.py --steps 1 --warmup 0 --nodes 1000000 --trees 0 --volume-trees 0 --cpu-sample 1000 > /dev/null 2> $OUT/r2r_ncu_machine_err.log
echo "exit $?"
echo "== R3 launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/r02r_launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --nodes 300000 --trees 0 --volume-trees 0 --cpu-sample 1000 > $OUT/r2r_launch_bench.json 2> $OUT/r2r_launch_err.log
echo "exit $?"; wc -l $OUT/r02r_launches_bench.csv
ls -la $OUT | tail -8
