#!/bin/bash
for f in build/variants/*.so; do
  echo "== $f"
  for n in 8192 3355 1638; do TVL1_SO=$f python scripts/kbench.py iterate $n 2>&1 | tail -1; done
done
