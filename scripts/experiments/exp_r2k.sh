echo "== base"; python scripts/quick_bench.py 8192:6 2>&1 | grep -E "rep1"
for v in u0t0r0 u0t1r0 u0t0r1 u0t1r1; do
  echo "== $v"
  TVL1_SO=build/variants/$v.so python scripts/quick_bench.py 8192:6 2>&1 | grep -E "rep1"
done
echo "== masked 20% bands: base, t1r0, t1r1"
MASK_FRAC=0.2 python scripts/quick_bench.py 4096:5 2>&1 | grep -E "rep1"
MASK_FRAC=0.2 TVL1_SO=build/variants/u0t1r0.so python scripts/quick_bench.py 4096:5 2>&1 | grep -E "rep1"
MASK_FRAC=0.2 TVL1_SO=build/variants/u0t1r1.so python scripts/quick_bench.py 4096:5 2>&1 | grep -E "rep1"
TVL1_SO=build/variants/u0t1r1.so python -m pytest tests/test_gpu_solve.py tests/test_gpu_kernels.py tests/test_gpu_random_configs.py tests/test_gpu_arith.py -m gpu -x -q 2>&1 | tail -5
