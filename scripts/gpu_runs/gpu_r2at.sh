#!/bin/bash
# round-2 GPU call AT (2 GPUs): bench at N=2 as the driver launches it
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/r2at_bench_n2.json 2> $OUT/r2at_bench_n2_err.log; echo "exit $?"; cut -c1-300 $OUT/r2at_bench_n2.json; tail -3 $OUT/r2at_bench_n2_err.log
