python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solve.py tests/test_gpu_random_configs.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -3
echo "== packed warp";  python scripts/quick_bench.py 8192:6 4096:5 2>&1 | grep -E "rep1"
python scripts/kbench.py warp 8192 2
echo "== scalar warp"; TVL1_SO=build/variants/warp_scalar.so python scripts/quick_bench.py 8192:6 4096:5 2>&1 | grep -E "rep1"
TVL1_SO=build/variants/warp_scalar.so python scripts/kbench.py warp 8192 2
