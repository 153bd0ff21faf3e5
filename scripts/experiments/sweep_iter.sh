#!/bin/bash
# developer A/B: times k_iterate / k_iterate2 of every library under build/variants/
for f in build/variants/*.so; do
  echo "== $f"
  for n in ${SIZES:-8192 3355 2684}; do TVL1_SO=$f python scripts/kbench.py iterate $n 2>&1 | tail -2; done
done
