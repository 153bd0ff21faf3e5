# r2o: k_warp occupancy / tile variants after the packed sums
echo "== base"; python scripts/kbench.py warp 8192 2; python scripts/quick_bench.py 8192:6 4096:5 2048:5 2>&1 | grep rep1
for v in minb3 minb2; do echo "== $v"; TVL1_SO=build/variants/warp_$v.so python scripts/kbench.py warp 8192 2; TVL1_SO=build/variants/warp_$v.so python scripts/quick_bench.py 8192:6 4096:5 2048:5 2>&1 | grep rep1; done
