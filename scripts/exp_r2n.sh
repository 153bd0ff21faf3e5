python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solve.py tests/test_gpu_random_configs.py tests/test_gpu_cli.py tests/test_gpu_stack.py -m gpu -x -q 2>&1 | tail -4
