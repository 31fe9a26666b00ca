#!/bin/bash
# One gpurun call: parity tests, bench lines, ncu launch list + one full capture of a bulk machine_kernel slice, knob runs.
# usage (under gpurun): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01e}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/${TAG}_smoke.log
echo "== bench"; timeout 600 python bench.py > $OUT/${TAG}_bench_line.json 2> $OUT/${TAG}_bench_err.log; echo "bench exit $?"; cut -c1-400 $OUT/${TAG}_bench_line.json
echo "== bench reference"; timeout 300 python bench.py --impl reference > $OUT/${TAG}_bench_reference_line.json 2>> $OUT/${TAG}_bench_err.log; echo "ref exit $?"; cut -c1-300 $OUT/${TAG}_bench_reference_line.json
echo "== hybrid slice log"; GLC_SLICE_LOG=1 timeout 300 python scripts/knobs.py 1000000 > $OUT/${TAG}_knobs_default.log 2>&1; tail -25 $OUT/${TAG}_knobs_default.log
for K in "GLC_DRAIN_BELOW=30000" "GLC_DRAIN_BELOW=120000" "GLC_DRAIN_DENSE_BUDGET=128" "GLC_DRAIN_DENSE_BUDGET=1024" "GLC_DRAIN_EXPRESS=0" "LIB=scripts/_variants/libglcb200_s1024.so" "LIB=scripts/_variants/libglcb200_s1024.so GLC_DRAIN_BELOW=30000"; do
  F=$OUT/${TAG}_knobs_$(echo "$K" | tr ' /=' '___').log
  GLC_SLICE_LOG=1 timeout 300 python scripts/knobs.py 1000000 $K > $F 2>&1; grep KNOBS $F
done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 1 --warmup 1 --nodes 300000 --cpu-sample 20000 > $OUT/${TAG}_launches_bench.log 2>&1; echo "ncu list exit $?"
echo "== ncu full: second machine_kernel slice (all slots busy)"
GLC_SLICE_BUDGET=4096 GLC_MAX_SLICES=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:machine_kernel -s 1 -c 1 -o $OUT/${TAG}_machine_bulk -f python scripts/prof_slices.py 1000000 > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full exit $?"; tail -2 $OUT/${TAG}_ncu_full.log
ls -la $OUT | tail -30
