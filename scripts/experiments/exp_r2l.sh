echo "== base"; python scripts/quick_bench.py 8192:6 2>&1 | grep -E "rep1"
for v in t1 t1n1; do
  echo "== $v"
  TVL1_SO=build/variants/$v.so python scripts/quick_bench.py 8192:6 2>&1 | grep -E "rep1"
  MASK_FRAC=0.2 TVL1_SO=build/variants/$v.so python scripts/quick_bench.py 4096:5 2>&1 | grep -E "rep1"
done
TVL1_SO=build/variants/t1n1.so python -m pytest tests/test_gpu_solve.py tests/test_gpu_kernels.py tests/test_gpu_random_configs.py tests/test_gpu_arith.py -m gpu -x -q 2>&1 | tail -5
