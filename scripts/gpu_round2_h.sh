#!/bin/bash
# round-2 GPU call H: suites + bench (1 GPU)
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_round2.py -q -m gpu 2>&1 | tail -60 > $OUT/r02h_tests_new.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -60 > $OUT/r02h_tests_old.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/r02h_bench.json 2> $OUT/r02h_bench.err
tail -3 $OUT/r02h_tests_new.log; tail -3 $OUT/r02h_tests_old.log; tail -c 300 $OUT/r02h_bench.err
