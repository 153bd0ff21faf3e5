echo "== base"; python scripts/kbench.py outer 8192 2; python scripts/quick_bench.py 8192:6 2>&1 | grep -E "rep1"
echo "== park"; TVL1_SO=build/variants/park.so python scripts/kbench.py outer 8192 2; TVL1_SO=build/variants/park.so python scripts/quick_bench.py 8192:6 2>&1 | grep -E "rep1"
TVL1_SO=build/variants/park.so python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fused2 or outer" 2>&1 | tail -2
