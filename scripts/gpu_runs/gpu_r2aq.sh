#!/bin/bash
# round-2 GPU call AQ: ncu --set full of the express drain launch (592 predicted-longest nodes, one per warp, to completion) with the
# cooperative GK15 pass
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:drain_kernel -s 0 -c 1 -f -o $OUT/r02aq_drain_express \
  python bench.py --steps 1 --warmup 0 --nodes 1000000 --trees 0 --volume-trees 0 --cpu-sample 1000 > /dev/null 2> $OUT/r2aq_ncu_err.log
echo "ncu exit $?"; ls -la $OUT | grep r02aq
