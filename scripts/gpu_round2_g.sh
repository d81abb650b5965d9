#!/bin/bash
# round-2 GPU call G: ncu over one launch of every aisp kernel (report kept on the box, text summary
# brought back), compute-sanitizer over every kernel at ragged sizes
set -u
OUT=gpurun_out
mkdir -p $OUT
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section SchedulerStats --section WarpStateStats --section ComputeWorkloadAnalysis"
timeout 900 ncu $SEC --clock-control none --kernel-name-base demangled -k regex:aisp:: -f -o /tmp/prof_all_r02 \
    python scripts/profile_all_kernels.py > $OUT/r02g_ncu_all.log 2>&1
python scripts/ncu_summary.py /tmp/prof_all_r02.ncu-rep > $OUT/r02_all_kernels_ncu_summary.txt 2>&1
grep "section:" $OUT/r02g_ncu_all.log > $OUT/r02_all_kernels_sections.txt
ls -la /tmp/prof_all_r02.ncu-rep >> $OUT/r02g_ncu_all.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanity_all_kernels.py > $OUT/r02g_memcheck.log 2>&1
echo "memcheck exit $?" >> $OUT/r02g_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanity_all_kernels.py > $OUT/r02g_racecheck.log 2>&1
echo "racecheck exit $?" >> $OUT/r02g_racecheck.log
wc -l $OUT/r02_all_kernels_ncu_summary.txt; tail -3 $OUT/r02g_memcheck.log; tail -3 $OUT/r02g_racecheck.log; du -sh $OUT
