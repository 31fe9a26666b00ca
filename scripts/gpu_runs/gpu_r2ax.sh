#!/bin/bash
# round-2 GPU call AX: final build (stream_spread = 2): full GPU suite, both bench arms
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r2ax_pytest_gpu.log 2>&1; echo "exit $?"; tail -3 $OUT/r2ax_pytest_gpu.log
timeout 1200 python bench.py > $OUT/r2ax_bench_line.json 2> $OUT/r2ax_bench_err.log; echo "exit $?"; cut -c1-260 $OUT/r2ax_bench_line.json
timeout 1200 python bench.py --impl reference > $OUT/r2ax_bench_reference_line.json 2> $OUT/r2ax_bench_ref_err.log; echo "exit $?"; cut -c1-200 $OUT/r2ax_bench_reference_line.json
