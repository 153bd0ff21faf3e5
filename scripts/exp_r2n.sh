python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solve.py tests/test_gpu_random_configs.py -m gpu -x -q 2>&1 | tail -3
echo "== chunked";  python scripts/quick_bench.py 8192:6 2>&1 | grep -E "rep1|L0|L1|L3"
python scripts/kbench.py outer 8192 2
echo "== tiles"; TVL1_SO=build/variants/tiles.so python scripts/quick_bench.py 8192:6 2>&1 | grep -E "rep1|L0|L1|L3"
TVL1_SO=build/variants/tiles.so python scripts/kbench.py outer 8192 2
echo "== chunked 4096 / 2048 / roi"; python scripts/quick_bench.py 4096:5 2048:5 2>&1 | grep -E "rep1"
echo "== tiles 4096 / 2048"; TVL1_SO=build/variants/tiles.so python scripts/quick_bench.py 4096:5 2048:5 2>&1 | grep -E "rep1"
