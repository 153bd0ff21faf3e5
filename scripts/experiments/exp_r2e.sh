python -m pytest tests -m gpu -x -q -k "not 16384 and not 8192 and not 6144" 2>&1 | tail -4
python scripts/kbench.py all 8192 3
TVL1_SO=build/variants/med4.so python scripts/kbench.py median 8192 5
python scripts/kbench.py warp 2684 5
python scripts/quick_bench.py 8192:6 2>&1 | grep -E "rep1"
python scripts/quick_bench.py 4096:5 2>&1 | grep -E "rep1"
python scripts/quick_bench.py 2048:5 2>&1 | grep -E "rep1"
