#!/bin/bash
# developer sweep: forced tile heights for both iteration kernels
for n in ${SIZES:-8192 2684}; do
for r in ${ROWS:-8 16 24 32 48 64}; do
  echo "== n=$n R=$r"; TVL1_DEV_VERBOSE=1 TVL1_DEV_ROWS=$r TVL1_DEV_ROWS2=$r python scripts/kbench.py iterate $n 2>&1 | grep -E "k_iterate|tile_rows" | sort | uniq | tail -4
done
echo "== n=$n auto"; TVL1_DEV_VERBOSE=1 python scripts/kbench.py iterate $n 2>&1 | grep -E "k_iterate|tile_rows" | sort | uniq | tail -4
done
