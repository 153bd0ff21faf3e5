for v in ws4; do
  echo "== $v"
  TVL1_SO=build/variants/$v.so TVL1_DEV_VERBOSE=1 python scripts/kbench.py outer 8192 3 2>&1 | tail -3
  TVL1_SO=build/variants/$v.so python scripts/quick_bench.py 8192:6 2048:5 2>&1 | grep -E "rep1|L0|L5"
  TVL1_SO=build/variants/$v.so python scripts/roi_probe.py 2>&1 | tail -4
done
TVL1_SO=build/variants/ws4.so python -m pytest tests/test_gpu_solve.py tests/test_gpu_kernels.py tests/test_gpu_random_configs.py -m gpu -x -q 2>&1 | tail -5
