#!/bin/bash
# round-2 GPU call C: suites, bench (1 GPU), NLM ncu
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_round2.py -q -m gpu 2>&1 | tail -200 > $OUT/r02c_tests_new.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -120 > $OUT/r02c_tests_old.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/r02c_bench.json 2> $OUT/r02c_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nlm_kernel -c 1 -f -o $OUT/prof_nlm_r02c \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > $OUT/r02c_ncu_nlm.log 2>&1
tail -4 $OUT/r02c_tests_new.log; tail -4 $OUT/r02c_tests_old.log; tail -c 600 $OUT/r02c_bench.err
