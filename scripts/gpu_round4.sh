#!/bin/bash
# validation of the 256-thread machine build: parity, smoke, bench lines, launch list
TAG=${1:-r01h}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest_gpu.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/${TAG}_smoke.log
echo "== bench"; timeout 600 python bench.py > $OUT/${TAG}_bench_line.json 2> $OUT/${TAG}_bench_err.log; echo "bench exit $?"; cut -c1-300 $OUT/${TAG}_bench_line.json; tail -3 $OUT/${TAG}_bench_err.log
echo "== bench reference"; timeout 300 python bench.py --impl reference > $OUT/${TAG}_bench_reference_line.json 2>> $OUT/${TAG}_bench_err.log; echo "ref exit $?"; cut -c1-200 $OUT/${TAG}_bench_reference_line.json
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 1 --warmup 1 --nodes 300000 --cpu-sample 20000 --trees 0 > $OUT/${TAG}_launches_bench.log 2>&1; echo "ncu list exit $?"
