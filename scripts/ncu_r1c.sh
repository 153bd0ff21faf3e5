#!/bin/bash
# ncu passes for profiles/r1c (run under gpurun): launch list of the bench command, then full-set
# captures of one LEVEL-0 launch of each heavy kernel.  A plain run precedes every ncu run
# (B200_PROFILING.md); numbers printed under ncu are never bench values.
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --stack-pairs 0"
T=${TAG:-r1c}
$CMD > gpurun_out/plain_$T.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/launches_$T.csv $CMD > gpurun_out/ncu_list_$T.log 2>&1
# per kernel: index (among that kernel's own launches) of the longest launch of the 4th pair
python - "$T" > gpurun_out/skips_$T.txt <<'PY'
import csv, sys, re
T = sys.argv[1]
rows = []
with open("gpurun_out/launches_%s.csv" % T) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", ""))))
starts = [i for i, (k, _) in enumerate(rows) if "k_convert_u8" in k][0::2]
lo = starts[3] if len(starts) > 3 else starts[-1]
hi = starts[4] if len(starts) > 4 else len(rows)
for pat in ("k_iterate2", "k_iterate<", "k_median5", "k_warp"):
    idx = [i for i, (k, _) in enumerate(rows) if pat in k]
    inpair = [i for i in idx if lo <= i < hi]
    if not inpair:
        continue
    best = max(inpair, key=lambda i: rows[i][1])
    print(pat.rstrip("<"), idx.index(best))
PY
cat gpurun_out/skips_$T.txt
while read name skip; do
  $CMD > gpurun_out/plain_${T}_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"${name}" -s $skip -c 1 \
      -o gpurun_out/prof_${name}_$T -f $CMD > gpurun_out/ncu_${name}_$T.log 2>&1
done < gpurun_out/skips_$T.txt
ls -la gpurun_out/ | tail -8
