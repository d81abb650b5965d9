#!/bin/bash
# round-2 final evidence call (1 GPU): tests, bench line, launch list of the bench command, one launch of every kernel
# under ncu (section set), --set full captures of the kernels the round worked on (text summaries only: reports stay on
# the box), compute-sanitizer memcheck / racecheck.
set -u
OUT=gpurun_out
mkdir -p $OUT
T=r02z
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > $OUT/${T}_tests.log; tail -2 $OUT/${T}_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err; tail -c 200 $OUT/${T}_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${T}_smoke.log 2>&1; tail -2 $OUT/${T}_smoke.log
timeout 600 python scripts/micro/laggard_bench.py > $OUT/${T}_lag.jsonl 2>&1
timeout 600 python scripts/micro/chain_bench.py > $OUT/${T}_chain.jsonl 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_${T}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > $OUT/launches_${T}.log 2>&1
python scripts/launch_shares.py $OUT/launches_${T}.csv > $OUT/${T}_launch_shares.txt 2>&1; head -8 $OUT/${T}_launch_shares.txt
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats --section ComputeWorkloadAnalysis"
timeout 900 ncu $SEC --clock-control none --kernel-name-base demangled -k regex:aisp:: -f -o /tmp/prof_all \
    python scripts/profile_all_kernels.py > $OUT/${T}_ncu_all.log 2>&1
python scripts/ncu_summary.py /tmp/prof_all.ncu-rep > $OUT/${T}_all_kernels_ncu_summary.txt 2>&1
grep "section:" $OUT/${T}_ncu_all.log > $OUT/${T}_all_kernels_sections.txt
full() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o /tmp/prof_$name "$@" > $OUT/${T}_ncu_${name}.log 2>&1
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > /tmp/raw_$name.csv 2>/dev/null
  ncu -i /tmp/prof_$name.ncu-rep --page source --csv > /tmp/src_$name.csv 2>/dev/null
  python scripts/ncu_raw_summary.py /tmp/raw_$name.csv > $OUT/${T}_full_${name}.txt 2>&1
  for k in $(echo "$rx" | tr '|' ' '); do python scripts/ncu_opcode_hist.py /tmp/src_$name.csv $k >> $OUT/${T}_full_${name}.txt 2>&1; done
}
full nlm2 nlm2_kernel 3 1 python scripts/micro/nlm_ab.py
full shr512 "sharpen_kernel|sharpen_adjoint" 3 7 python scripts/micro/laggard_bench.py --iters 1 --only shr512
full usm4k "sharpen_kernel|sharpen_adjoint" 3 7 python scripts/micro/laggard_bench.py --iters 1 --only usm4k
full chain pw_chain_fixed 4 1 python scripts/micro/chain_bench.py --iters 2 --cases 1
full bank "pw_bank_fwd|pw_bank_bwd" 2 2 python bench.py --steps 2 --warmup 3 --no-cpu --no-extras
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanity_all_kernels.py > $OUT/${T}_memcheck.log 2>&1
echo "memcheck exit $?" >> $OUT/${T}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanity_all_kernels.py > $OUT/${T}_racecheck.log 2>&1
echo "racecheck exit $?" >> $OUT/${T}_racecheck.log
tail -3 $OUT/${T}_memcheck.log; tail -3 $OUT/${T}_racecheck.log; du -sh $OUT
