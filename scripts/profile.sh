#!/bin/bash
# ncu evidence for one round: launch list of the bench step + full captures of the top kernels.
# Run under gpurun (1 GPU):  bash scripts/profile.sh r01
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/launches_$TAG.log 2>&1
KERNELS=${KERNELS:-"nlm_kernel sharpen_kernel pw_bank_fwd_kernel pw_bank_bwd_kernel"}
[ "${PERFILTER:-0}" = 1 ] && KERNELS="nlm_kernel sharpen_kernel pw_fwd_kernel pw_bwd_kernel"
for K in $KERNELS; do
  # the first launches of every kernel are the banked ones (all slots of a family in one launch);
  # PERFILTER=1 skips past warm-up + timed steps to the one-filter-at-a-time launches of the breakdown
  N=2; S=0
  if [ "${PERFILTER:-0}" = 1 ]; then S=5; [ $K = pw_fwd_kernel ] && N=8; [ $K = pw_bwd_kernel ] && N=8; fi
  ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $N -o $OUT/prof_${K}_$TAG -f $BENCH > $OUT/prof_${K}_$TAG.log 2>&1
done
ls -la $OUT
for K in $KERNELS; do
  ncu -i $OUT/prof_${K}_$TAG.ncu-rep --page details > $OUT/details_${K}_$TAG.txt 2>&1
done
