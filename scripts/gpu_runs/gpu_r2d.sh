#!/bin/bash
# round-2 GPU call D: scale regression test on the new and the old queue protocol, full GPU suite, smoke, full bench
OUT=gpurun_out
mkdir -p $OUT
echo "== D1 scale tests, product build"
timeout 900 python -m pytest tests/test_gpu_machine_scale.py -m gpu -x -q > $OUT/r2d_scale_new.log 2>&1; echo "exit $?"; tail -3 $OUT/r2d_scale_new.log
echo "== D2 scale tests, round-1 queue protocol (expected to FAIL: this is the reproducer)"
GLC_LIB_PATH=$PWD/galacticus_b200/libglcb200_oldq.so timeout 600 python -m pytest tests/test_gpu_machine_scale.py -m gpu -q > $OUT/r2d_scale_oldq.log 2>&1; echo "exit $?"; tail -8 $OUT/r2d_scale_oldq.log | cut -c1-300
echo "== D3 full GPU suite"
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2d_pytest_gpu.log 2>&1; echo "exit $?"; tail -3 $OUT/r2d_pytest_gpu.log
echo "== D4 smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2d_smoke.log 2>&1; echo "exit $?"; tail -3 $OUT/r2d_smoke.log
echo "== D5 bench (default flags)"
timeout 900 python bench.py > $OUT/r2d_bench_line.json 2> $OUT/r2d_bench_err.log; echo "bench exit $?"; cut -c1-400 $OUT/r2d_bench_line.json; tail -3 $OUT/r2d_bench_err.log
echo "== D6 reference arm"
timeout 600 python bench.py --impl reference > $OUT/r2d_bench_reference_line.json 2> $OUT/r2d_bench_ref_err.log; echo "exit $?"; cut -c1-300 $OUT/r2d_bench_reference_line.json
