#!/bin/bash
# round-2 GPU call AC: ncu --set full of the DENSE drain pass alone (express kernel off so that nothing overlaps it)
OUT=gpurun_out
mkdir -p $OUT
GLC_DRAIN_EXPRESS=0 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:drain_kernel -s 0 -c 1 -f -o $OUT/r02ac_drain_dense \
  python bench.py --steps 1 --warmup 0 --nodes 1000000 --trees 0 --volume-trees 0 --cpu-sample 1000 > /dev/null 2> $OUT/r2ac_ncu_err.log
echo "exit $?"; ls -la $OUT | tail -3
