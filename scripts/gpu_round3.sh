#!/bin/bash
# final GPU call of the round: parity with the new defaults, smoke, bench lines, launch list, thread-count variants, forest scaling
TAG=${1:-r01g}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/${TAG}_smoke.log
echo "== bench"; timeout 900 python bench.py > $OUT/${TAG}_bench_line.json 2> $OUT/${TAG}_bench_err.log; echo "bench exit $?"; cut -c1-300 $OUT/${TAG}_bench_line.json; tail -3 $OUT/${TAG}_bench_err.log
echo "== bench reference"; timeout 600 python bench.py --impl reference > $OUT/${TAG}_bench_reference_line.json 2>> $OUT/${TAG}_bench_err.log; echo "ref exit $?"; cut -c1-200 $OUT/${TAG}_bench_reference_line.json
echo "== knobs / variants"
GLC_SLICE_LOG=1 timeout 300 python scripts/knobs.py 1000000 > $OUT/${TAG}_knobs_default.log 2>&1; grep KNOBS $OUT/${TAG}_knobs_default.log
for K in "LIB=scripts/_variants/libglcb200_t384.so" "LIB=scripts/_variants/libglcb200_t256.so"; do
  F=$OUT/${TAG}_knobs_$(echo "$K" | tr ' /=' '___').log
  GLC_SLICE_LOG=1 timeout 300 python scripts/knobs.py 1000000 $K > $F 2>&1; grep KNOBS $F
done
echo "== forest scaling"
timeout 900 python scripts/forest_bench.py 4000 0 2>&1 | grep FOREST | cut -c1-700 | tee $OUT/${TAG}_forest_4000.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 1 --warmup 1 --nodes 300000 --cpu-sample 20000 --trees 0 > $OUT/${TAG}_launches_bench.log 2>&1; echo "ncu list exit $?"
ls $OUT | wc -l
