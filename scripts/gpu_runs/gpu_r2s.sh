#!/bin/bash
# round-2 GPU call S: validation of the final build -- full GPU suite, smoke, both bench arms
OUT=gpurun_out
mkdir -p $OUT
echo "== S1 full GPU suite"
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r2s_pytest_gpu.log 2>&1; echo "exit $?"; tail -3 $OUT/r2s_pytest_gpu.log
echo "== S2 smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2s_smoke.log 2>&1; echo "exit $?"; tail -3 $OUT/r2s_smoke.log
echo "== S3 bench"
timeout 1200 python bench.py > $OUT/r2s_bench_line.json 2> $OUT/r2s_bench_err.log; echo "exit $?"; cut -c1-300 $OUT/r2s_bench_line.json; tail -3 $OUT/r2s_bench_err.log
echo "== S4 bench --impl reference"
timeout 1200 python bench.py --impl reference > $OUT/r2s_bench_reference_line.json 2> $OUT/r2s_bench_ref_err.log; echo "exit $?"; cut -c1-600 $OUT/r2s_bench_reference_line.json
