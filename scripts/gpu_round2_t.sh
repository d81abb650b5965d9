#!/bin/bash
# quick GPU call: both parity suites
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_round2.py -q -m gpu 2>&1 | tail -80 > $OUT/r02t_tests_new.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -40 > $OUT/r02t_tests_old.log
tail -3 $OUT/r02t_tests_new.log; tail -3 $OUT/r02t_tests_old.log
