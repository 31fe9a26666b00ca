#!/bin/bash
# round-2 GPU call AU: 320 threads per machine block (10 warps, 168 registers, 44 bytes of spills) against 256 (8 warps, 198 registers)
for lib in libglcb200_m320.so libglcb200.so libglcb200_m320.so libglcb200.so; do
  GLC_LIB_PATH=$PWD/galacticus_b200/$lib timeout 300 python scripts/knobs.py 1000000 GLC_SLICE_LOG=1 2> gpurun_out/r2au.err | grep KNOBS
  grep "(hold)" gpurun_out/r2au.err | tail -1 | cut -c1-100; echo $lib
done
