#!/bin/bash
# Final-round evidence (run under gpurun): the launch list of the BENCH command itself (its timed
# step: the 4th solve), after a plain run of the same command; and a full-set capture of the level-0
# median launch of that step.  Numbers printed under ncu are never bench values.
T=${TAG:-r1f}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --stack-pairs 0"
$CMD > gpurun_out/plain_$T.json 2> gpurun_out/plain_$T.err || exit 1
N=$(python -c "import json;print(json.loads(open('gpurun_out/plain_$T.json').read().strip().splitlines()[-1])['gpu_launches'])")
echo "launches per step: $N"
ncu --metrics gpu__time_duration.sum --clock-control none -s $((3 * N)) -c $N --csv \
    --log-file gpurun_out/launches_$T.csv $CMD > gpurun_out/ncu_list_$T.log 2>&1
grep -c k_iterate2 gpurun_out/launches_$T.csv
ncu --set full --clock-control none --import-source on -k regex:k_median5 -s 131 -c 1 \
    -o gpurun_out/prof_k_median5_$T -f $CMD > gpurun_out/ncu_median_$T.log 2>&1
ls -la gpurun_out/*_$T* | tail -6
