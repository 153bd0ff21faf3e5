for p in 0 1 2 3; do echo "== L2 promotion $p"; TVL1_DEV_L2PROMO=$p python scripts/kbench.py outer 8192 2; done
echo "== ring 4 (TMA)"; TVL1_SO=build/variants/tma4.so python scripts/kbench.py outer 8192 2
TVL1_SO=build/variants/tma4.so python scripts/quick_bench.py 8192:6 2>&1 | grep rep1
echo "== promo 3 pair"; TVL1_DEV_L2PROMO=3 python scripts/quick_bench.py 8192:6 2>&1 | grep rep1
echo "== base pair"; python scripts/quick_bench.py 8192:6 2>&1 | grep rep1
