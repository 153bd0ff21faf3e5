# r2o: k_warp occupancy / tile variants after the packed sums
echo "== base"; python scripts/kbench.py warp 8192 2
for v in minb3 th4; do echo "== $v"; TVL1_SO=build/variants/warp_$v.so python scripts/kbench.py warp 8192 2; done
