#!/bin/bash
# round-2 GPU call D (2 GPUs): bench under torchrun -- rollout + NCCL all-reduce section, PCIe probe with 2 ranks
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/r02d_bench_2gpu.json 2> $OUT/r02d_bench_2gpu.err
tail -c 1500 $OUT/r02d_bench_2gpu.err; tail -c 400 $OUT/r02d_bench_2gpu.json
