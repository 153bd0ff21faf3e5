#!/bin/bash
# ncu passes of round 2 (run under gpurun): TAG=r2a bash scripts/ncu_r2.sh
# 1. launch list of one timed pair of the bench workload (plain run first; numbers printed under ncu
#    are never bench values)
# 2. --set full capture of ONE k_outer launch (the shipped iteration kernel) at 8192^2 through the
#    stage-level entry point: KB_OUTER_ITERS iterations in the launch, so DRAM bytes per px-iteration
#    are bytes / (8192^2 * KB_OUTER_ITERS)
# 3. --set full of the level-0 k_median5 and k_warp launches (kbench)
T=${TAG:-r2a}
export KB_OUTER_ITERS=${KB_OUTER_ITERS:-10}
BCMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-parity --stack-pairs 0 --volume-pairs 0"
$BCMD > gpurun_out/plain_$T.log 2>&1 || { tail -5 gpurun_out/plain_$T.log; exit 1; }
N=$(python -c "import json,sys; print(json.loads(open('gpurun_out/plain_$T.log').read().strip().splitlines()[-1])['gpu_launches'])")
echo "launches per pair: $N"
# the bench solves 3 warm-up pairs + 1 timed pair (then the e2e pairs): skip the warm-up pairs
ncu --metrics gpu__time_duration.sum --clock-control none -s $((3 * N)) -c $N --csv \
    --log-file gpurun_out/launches_$T.csv $BCMD > gpurun_out/ncu_list_$T.log 2>&1
KCMD="python scripts/kbench.py outer 8192 2"
$KCMD > gpurun_out/plain_kb_$T.log 2>&1 || { tail -5 gpurun_out/plain_kb_$T.log; exit 1; }
cat gpurun_out/plain_kb_$T.log
ncu --set full --clock-control none --import-source on -k regex:k_outer -s 1 -c 1 \
    -o gpurun_out/prof_k_outer_$T -f $KCMD > gpurun_out/ncu_k_outer_$T.log 2>&1
if [ -z "$ONLY_OUTER" ]; then
KCMD2="python scripts/kbench.py median 8192 2"
$KCMD2 > gpurun_out/plain_kb2_$T.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_median5 -s 1 -c 1 \
    -o gpurun_out/prof_k_median5_$T -f $KCMD2 > gpurun_out/ncu_k_median5_$T.log 2>&1
KCMD3="python scripts/kbench.py warp 8192 2"
$KCMD3 > gpurun_out/plain_kb3_$T.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_warp -s 1 -c 1 \
    -o gpurun_out/prof_k_warp_$T -f $KCMD3 > gpurun_out/ncu_k_warp_$T.log 2>&1
fi
ls -la gpurun_out/*_$T* | tail -12
