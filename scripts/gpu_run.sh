#!/bin/bash
# one GPU call: scripts/gpu_run.sh TAG [tests] [bench] [nlm_ab] ...  (logs land in gpurun_out/TAG_*)
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for what in "$@"; do
  case $what in
    tests)
      timeout 1200 python -m pytest tests/test_gpu_round2.py -q -m gpu -x 2>&1 | tail -60 > $OUT/${TAG}_tests_new.log
      timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -60 > $OUT/${TAG}_tests_old.log
      tail -3 $OUT/${TAG}_tests_new.log; tail -3 $OUT/${TAG}_tests_old.log;;
    nlmtests)
      timeout 1200 python -m pytest tests -q -m gpu -x -k "nlm or NLM or denoise or bank or smoke or config" 2>&1 | tail -60 > $OUT/${TAG}_tests_nlm.log
      tail -5 $OUT/${TAG}_tests_nlm.log;;
    bench)
      timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
      tail -c 300 $OUT/${TAG}_bench.err;;
    nlm_ab)
      AISP_NLM_LAYOUT=1col timeout 300 python scripts/micro/nlm_ab.py 2>&1 | tee $OUT/${TAG}_nlm_ab.log
      timeout 300 python scripts/micro/nlm_ab.py 2>&1 | tee -a $OUT/${TAG}_nlm_ab.log;;
    ncu_nlm)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:nlm2_kernel -s 3 -c 1 -f -o $OUT/prof_nlm2_${TAG} \
          python scripts/micro/nlm_ab.py > $OUT/${TAG}_ncu_nlm.log 2>&1
      tail -3 $OUT/${TAG}_ncu_nlm.log;;
    chain)
      timeout 600 python scripts/micro/chain_bench.py 2>&1 | tee $OUT/${TAG}_chain.jsonl;;
    chain_ab)
      rm -f $OUT/${TAG}_chain_ab.jsonl
      for c in 31 32 21 22; do echo "cfg $c" | tee -a $OUT/${TAG}_chain_ab.jsonl
        AISP_CHAIN_CFG=$c timeout 600 python scripts/micro/chain_bench.py --cases 1 2>&1 | tee -a $OUT/${TAG}_chain_ab.jsonl; done;;
    ncu_chain)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_chain_fixed -s 4 -c 1 -f -o /tmp/prof_chainfixed \
          python scripts/micro/chain_bench.py --iters 2 --cases 1 > $OUT/${TAG}_ncu_chain.log 2>&1
      ncu -i /tmp/prof_chainfixed.ncu-rep --page details > $OUT/${TAG}_details_chain.txt 2>&1
      ncu -i /tmp/prof_chainfixed.ncu-rep --page raw --csv > $OUT/${TAG}_raw_chain.csv 2>&1
      ncu -i /tmp/prof_chainfixed.ncu-rep --page source --csv > $OUT/${TAG}_source_chain.csv 2>&1
      tail -3 $OUT/${TAG}_ncu_chain.log;;
    lag)
      timeout 600 python scripts/micro/laggard_bench.py 2>&1 | tee $OUT/${TAG}_lag.jsonl;;
    ncu_lag)
      # reports stay on the box (64 MiB pull limit): details / raw / source pages come back as text
      for c in ${NCU_CASES:-shr512 usm4k satp512}; do
        timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sharpen_kernel|sharpen_adjoint|pw_fwd|pw_bwd" -s ${NCU_SKIP:-3} -c ${NCU_COUNT:-6} -f \
            -o /tmp/prof_${c} python scripts/micro/laggard_bench.py --iters 1 --only $c > $OUT/${TAG}_ncu_${c}.log 2>&1
        ncu -i /tmp/prof_${c}.ncu-rep --page details > $OUT/${TAG}_details_${c}.txt 2>&1
        ncu -i /tmp/prof_${c}.ncu-rep --page raw --csv > $OUT/${TAG}_raw_${c}.csv 2>&1
        ncu -i /tmp/prof_${c}.ncu-rep --page source --csv > $OUT/${TAG}_source_${c}.csv 2>&1
        ls -la $OUT/${TAG}_*_${c}.*; done;;
    *) echo "unknown step $what";;
  esac
done
