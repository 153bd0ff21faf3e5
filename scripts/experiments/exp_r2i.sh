for v in r3m2; do
  echo "== $v"
  TVL1_SO=build/variants/$v.so TVL1_DEV_VERBOSE=1 python scripts/quick_bench.py 8192:6 2>&1 | grep -E "rep1|L0|L5|tile_rows 8192x8192"
done
for rmax in 80 96 128; do
  echo "== r3m3 rmax $rmax"
  TVL1_DEV_RMAX=$rmax TVL1_SO=build/variants/r3m3.so TVL1_DEV_VERBOSE=1 python scripts/quick_bench.py 8192:6 2>&1 | grep -E "rep1|L0|L5|tile_rows 8192x8192"
done
