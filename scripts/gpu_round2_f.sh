#!/bin/bash
# round-2 GPU call F: ncu --set full over one launch of every kernel (summary written on the box)
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "value_statistics or batched" 2>&1 | tail -5 > $OUT/r02f_tests.log
timeout 1500 ncu --set full --clock-control none -f -o $OUT/prof_all_r02 python scripts/profile_all_kernels.py > $OUT/r02f_ncu_all.log 2>&1
python scripts/ncu_summary.py $OUT/prof_all_r02.ncu-rep > $OUT/r02_all_kernels_ncu_summary.txt 2>&1
grep "section:" $OUT/r02f_ncu_all.log > $OUT/r02_all_kernels_sections.txt
ls -la $OUT/prof_all_r02.ncu-rep; tail -3 $OUT/r02f_tests.log; tail -3 $OUT/r02f_ncu_all.log; wc -l $OUT/r02_all_kernels_ncu_summary.txt
