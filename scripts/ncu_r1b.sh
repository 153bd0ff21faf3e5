#!/bin/bash
# ncu passes (run under gpurun): launch list of one whole pair + full-set captures of the level-0
# launches of the three heaviest kernels.  A plain run precedes each ncu run (B200_PROFILING.md).
set -x
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --stack-pairs 0"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 3318 -c 1106 --csv \
    --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_iterate -s 870 -c 3 \
    -o gpurun_out/prof_iterate_r1b -f $CMD > gpurun_out/ncu_iter.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_median5 -s 29 -c 1 \
    -o gpurun_out/prof_median_r1b -f $CMD > gpurun_out/ncu_med.log 2>&1
$CMD > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_warp -s 25 -c 1 \
    -o gpurun_out/prof_warp_r1b -f $CMD > gpurun_out/ncu_warp.log 2>&1
ls -la gpurun_out/ | tail -8
