# ring depth / blocks per SM of the two-iteration pass (developer experiment, run under gpurun)
for v in r4m3 r3m3 r3m4; do
  echo "== $v"
  TVL1_SO=build/variants/$v.so TVL1_DEV_VERBOSE=1 python scripts/kbench.py outer 8192 3 2>&1 | tail -4
  TVL1_SO=build/variants/$v.so python scripts/quick_bench.py 8192:6 2>&1 | grep -E "rep1|L0|L5"
done
TVL1_SO=build/variants/r3m4.so python -m pytest tests/test_gpu_solve.py tests/test_gpu_kernels.py tests/test_gpu_random_configs.py -m gpu -x -q 2>&1 | tail -3
