#!/bin/bash
# round-2 GPU call E: suites + bench + launch list (1 GPU)
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_round2.py -q -m gpu 2>&1 | tail -200 > $OUT/r02e_tests_new.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -120 > $OUT/r02e_tests_old.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/r02e_bench.json 2> $OUT/r02e_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_r02e.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > $OUT/launches_r02e.log 2>&1
tail -4 $OUT/r02e_tests_new.log; tail -4 $OUT/r02e_tests_old.log; tail -c 600 $OUT/r02e_bench.err
