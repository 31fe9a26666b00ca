#!/bin/bash
# round-2 GPU call BD: final build (no express launch at the hand-over): full GPU suite, both bench arms
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r2bd_pytest_gpu.log 2>&1; echo "exit $?"; tail -3 $OUT/r2bd_pytest_gpu.log
timeout 1200 python bench.py > $OUT/r2bd_bench_line.json 2> $OUT/r2bd_bench_err.log; echo "exit $?"; cut -c1-260 $OUT/r2bd_bench_line.json
timeout 1200 python bench.py --impl reference > $OUT/r2bd_bench_reference_line.json 2> $OUT/r2bd_bench_ref_err.log; echo "exit $?"; cut -c1-200 $OUT/r2bd_bench_reference_line.json
