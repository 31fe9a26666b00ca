#!/bin/bash
# round-2 GPU call AB: new GPU tests (edge-shaped forests, error surface, forests from XML documents)
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_forest_edge_cases.py tests/test_formats.py -m gpu -x -q > $OUT/r2ab_pytest.log 2>&1; echo "exit $?"; tail -15 $OUT/r2ab_pytest.log
