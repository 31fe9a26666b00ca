#!/bin/bash
# round-2 GPU call BA: both bench arms of the final build
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python bench.py > $OUT/r2ba_bench_line.json 2> $OUT/r2ba_bench_err.log; echo "exit $?"; cut -c1-260 $OUT/r2ba_bench_line.json
timeout 1200 python bench.py --impl reference > $OUT/r2ba_bench_reference_line.json 2> $OUT/r2ba_bench_ref_err.log; echo "exit $?"; cut -c1-200 $OUT/r2ba_bench_reference_line.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
