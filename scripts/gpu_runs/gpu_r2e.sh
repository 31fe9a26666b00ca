#!/bin/bash
# round-2 GPU call E: reproducer on the round-1 queue protocol; lone-lane ILP variants (QAG unroll, inlined elementary functions)
OUT=gpurun_out
mkdir -p $OUT
echo "== E0 scale tests on the round-1 queue protocol (expected to FAIL)"
GLC_LIB_PATH=$PWD/galacticus_b200/libglcb200_oldq.so timeout 600 python -m pytest tests/test_gpu_machine_scale.py -m gpu -q > $OUT/r2e_scale_oldq.log 2>&1; echo "exit $?"; tail -6 $OUT/r2e_scale_oldq.log | cut -c1-300
for v in base v1 v3 v2; do
  LIB=$PWD/galacticus_b200/libglcb200_$v.so
  [ $v = base ] && LIB=$PWD/galacticus_b200/libglcb200.so
  echo "== E-$v forest 1000"
  GLC_LIB_PATH=$LIB timeout 300 python scripts/forest_bench.py 1000 0 > $OUT/r2e_forest_$v.log 2>&1; grep FOREST $OUT/r2e_forest_$v.log | cut -c1-200
  echo "== E-$v node arm"
  GLC_LIB_PATH=$LIB timeout 300 python bench.py --trees 0 --volume-trees 0 --steps 2 --warmup 1 --cpu-sample 1000 > $OUT/r2e_bench_$v.json 2> $OUT/r2e_bench_$v.err; python - <<PY
import json
d=json.load(open("$OUT/r2e_bench_$v.json"))
k=d["roofline_kernels"]
print("$v", "ms_per_step %.1f"%d["ms_per_step"], {n:(round(k[n]["ms_per_pass"],1), round(k[n]["rhs_per_s"]/1e6,2)) for n in k})
PY
done
