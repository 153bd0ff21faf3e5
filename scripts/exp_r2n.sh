python -m pytest tests/test_gpu_solve.py tests/test_gpu_random_configs.py tests/test_gpu_stack.py tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -3
echo "== pz";  python scripts/quick_bench.py 8192:6 4096:5 2048:5 2>&1 | grep -E "rep1"
echo "== memset"; TVL1_SO=build/variants/memset.so python scripts/quick_bench.py 8192:6 4096:5 2048:5 2>&1 | grep -E "rep1"
