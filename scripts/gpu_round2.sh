#!/bin/bash
# second GPU call of the round: parity (incl. tree level), tree-level throughput, drain knobs, bench lines, FP64 op counts
TAG=${1:-r01f}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest_gpu.log
echo "== forest"
for K in "" "GLC_MACHINE_MIN_NODES=1000000000" "GLC_MACHINE_MIN_NODES=100000" "GLC_MACHINE_MIN_NODES=30000"; do
  timeout 600 python scripts/forest_bench.py 200 0 $K 2>&1 | grep FOREST | cut -c1-330
done
timeout 900 python scripts/forest_bench.py 1000 64 > $OUT/${TAG}_forest_1000.log 2>&1; grep FOREST $OUT/${TAG}_forest_1000.log | cut -c1-900
echo "== knobs"
for K in "GLC_DRAIN_BELOW=120000" "GLC_DRAIN_BELOW=160000" "GLC_DRAIN_BELOW=220000" "GLC_DRAIN_BELOW=120000 GLC_DRAIN_EXPRESS=0" "GLC_DRAIN_BELOW=160000 GLC_DRAIN_EXPRESS=0" "GLC_DRAIN_BELOW=120000 GLC_DRAIN_EXPRESS=0 GLC_DRAIN_DENSE_BUDGET=1024" "GLC_DRAIN_BELOW=120000 GLC_DRAIN_DENSE_BUDGET=1024" "GLC_DRAIN_BELOW=120000 GLC_DRAIN_EXPRESS=0 GLC_DRAIN_DENSE_BUDGET=192"; do
  F=$OUT/${TAG}_knobs_$(echo "$K" | tr ' /=' '___').log
  GLC_SLICE_LOG=1 timeout 300 python scripts/knobs.py 1000000 $K > $F 2>&1; grep KNOBS $F
done
echo "== bench"; timeout 600 python bench.py > $OUT/${TAG}_bench_line.json 2> $OUT/${TAG}_bench_err.log; echo "bench exit $?"; cut -c1-300 $OUT/${TAG}_bench_line.json
echo "== bench reference"; timeout 300 python bench.py --impl reference > $OUT/${TAG}_bench_reference_line.json 2>> $OUT/${TAG}_bench_err.log; echo "ref exit $?"; cut -c1-200 $OUT/${TAG}_bench_reference_line.json
echo "== FP64 op counts of one whole pass (300k nodes)"
timeout 900 ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,gpu__time_duration.sum --clock-control none -k regex:"machine_kernel|drain_kernel" --csv --log-file $OUT/${TAG}_fp64_ops.csv python scripts/prof_slices.py 300000 > $OUT/${TAG}_fp64_ops.log 2>&1; echo "ncu ops exit $?"; tail -1 $OUT/${TAG}_fp64_ops.log | cut -c1-300
ls $OUT | wc -l
