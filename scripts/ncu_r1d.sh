#!/bin/bash
# ncu passes for profiles/r1d (run under gpurun).  Plain run first (B200_PROFILING.md); numbers printed
# under ncu are never bench values.  The launch list covers ONE whole pair (quick_bench's second
# solve: launches 680..1359); the full-set captures take level-0 launches, found by index inside that
# pair (the level-0 launches are its last ones).
T=${TAG:-r1d}
CMD="python scripts/quick_bench.py 8192:6"
$CMD > gpurun_out/plain_$T.log 2>&1 || exit 1
N=$(grep -o "launches [0-9]*" gpurun_out/plain_$T.log | tail -1 | cut -d' ' -f2)
echo "launches per pair: $N"
ncu --metrics gpu__time_duration.sum --clock-control none -s $N -c $N --csv \
    --log-file gpurun_out/launches_$T.csv $CMD > gpurun_out/ncu_list_$T.log 2>&1
python - "$T" "$N" > gpurun_out/skips_$T.txt <<'PY'
import csv, sys
T, N = sys.argv[1], int(sys.argv[2])
lines = [l for l in open("gpurun_out/launches_%s.csv" % T) if not l.startswith("==")]
rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", ""))) for r in csv.DictReader(lines)
        if r.get("Metric Name") == "gpu__time_duration.sum"]
for pat in ("k_iterate2", "k_median5", "k_warp"):
    idx = [i for i, (k, _) in enumerate(rows) if pat in k]
    if not idx:
        continue
    best = max(idx, key=lambda i: rows[i][1])
    # launches of this kernel before `best` in this pair + all of them in the first pair
    print(pat.rstrip("<"), len(idx) + idx.index(best), pat)
PY
cat gpurun_out/skips_$T.txt
while read name skip pat; do
  rx="$name"; [ "$name" = "k_iterate" ] && rx="k_iterate<"
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c 1 \
      -o gpurun_out/prof_${name}_$T -f $CMD > gpurun_out/ncu_${name}_$T.log 2>&1
done < gpurun_out/skips_$T.txt
ls -la gpurun_out/*_$T* | tail -8
