#!/bin/bash
# round-2 GPU call A: new parity tests (all, no -x), chain timing, ncu of the chain kernel, old suite
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/r02a_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu 2>&1 | tail -150 > $OUT/r02a_tests_new.log
timeout 300 python scripts/micro/chain_bench.py --iters 20 > $OUT/r02a_chain_bench.jsonl 2> $OUT/r02a_chain_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pw_chain_bwd -c 2 -f -o $OUT/prof_chain_r02a \
    python scripts/micro/chain_bench.py --iters 1 > $OUT/r02a_ncu_chain.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -60 > $OUT/r02a_tests_old.log
timeout 120 python scripts/micro/pcie_bw.py > $OUT/r02a_pcie.log 2>&1
tail -5 $OUT/r02a_tests_new.log; cat $OUT/r02a_chain_bench.jsonl; tail -3 $OUT/r02a_tests_old.log
