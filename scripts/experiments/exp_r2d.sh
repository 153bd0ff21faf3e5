python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solve.py tests/test_gpu_random_configs.py tests/test_gpu_golden.py tests/test_gpu_stack.py -x -q 2>&1 | tail -3
python scripts/kbench.py median 8192 5
python scripts/kbench.py median 2684 5
TVL1_SO=build/variants/med3.so python -m pytest tests/test_gpu_kernels.py -x -q -k median 2>&1 | tail -2
TVL1_SO=build/variants/med3.so python scripts/kbench.py median 8192 5
TVL1_SO=build/variants/med3.so python scripts/kbench.py median 2684 5
python scripts/quick_bench.py 8192:6 2>&1 | grep -E "rep1"
python scripts/quick_bench.py 4096:5 2>&1 | grep -E "rep1"
