#!/bin/bash
# round-2 GPU call B: full parity suites (no -x), chain timing + ncu, one bench run
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_round2.py -q -m gpu 2>&1 | tail -200 > $OUT/r02b_tests_new.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -120 > $OUT/r02b_tests_old.log
timeout 300 python scripts/micro/chain_bench.py --iters 20 > $OUT/r02b_chain_bench.jsonl 2> $OUT/r02b_chain_bench.err
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/r02b_bench.json 2> $OUT/r02b_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pw_chain_bwd -c 1 -f -o $OUT/prof_chain_r02b \
    python scripts/micro/chain_bench.py --iters 1 > $OUT/r02b_ncu_chain.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02b_smoke.log 2>&1
tail -4 $OUT/r02b_tests_new.log; tail -4 $OUT/r02b_tests_old.log; cat $OUT/r02b_chain_bench.jsonl; tail -c 600 $OUT/r02b_bench.err; tail -2 $OUT/r02b_smoke.log
