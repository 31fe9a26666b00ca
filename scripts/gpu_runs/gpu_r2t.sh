#!/bin/bash
# round-2 GPU call T (2 GPUs): the N > 1 path of bench.py as the driver launches it
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 > $OUT/r2t_bench_2gpu.json 2> $OUT/r2t_bench_2gpu.err; echo "exit $?"
tail -c 1500 $OUT/r2t_bench_2gpu.json | head -c 1500; echo; tail -3 $OUT/r2t_bench_2gpu.err
